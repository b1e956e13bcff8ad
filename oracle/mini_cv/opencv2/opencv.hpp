// mini_cv — a minimal stand-in for the OpenCV C++ API.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: this container has no OpenCV C++ headers or libraries, so the reference's own
// translation units (/root/reference/src/Stabilizer.cpp, RollCorrection.cpp, AutoZoomCrop.cpp) cannot be
// compiled against the real thing.  This header declares exactly the slice of the cv:: API those files use.
// Containers (Mat, Vec, Point, Rect ...) are implemented here; every image-processing call
// (resize, cvtColor, goodFeaturesToTrack, calcOpticalFlowPyrLK, estimateAffinePartial2D, warpAffine,
// copyMakeBorder, KalmanFilter, Canny, HoughLines ...) is forwarded through a table of C callbacks
// (mini_cv_ops, see mini_cv_ops.h) that oracle/ref_lib.py fills with the REAL OpenCV 4.13 functions of the
// cv2 Python wheel.  So: reference host logic = the reference's own unmodified C++; OpenCV arithmetic = real
// OpenCV.  Nothing under video-stab_b200/ includes or links this.
#ifndef MINI_CV_OPENCV_HPP
#define MINI_CV_OPENCV_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mini_cv_ops.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_VERSION "mini_cv (ops: cv2 wheel)"

namespace cv {

inline int cvDepthSize(int depth) {
    static const int s[] = {1, 1, 2, 2, 4, 4, 8, 2};
    return s[depth & 7];
}
inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvRound(float v) { return (int)std::nearbyintf(v); }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) {
    int iv = cvRound(v);
    return (uchar)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
}
template <typename T> static inline T saturate_cast(float v) { return saturate_cast<T>((double)v); }
template <typename T> static inline T saturate_cast(int v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)(v < 0 ? 0 : v > 255 ? 255 : v); }

class Exception : public std::exception {
public:
    Exception() {}
    explicit Exception(const std::string &m) : msg(m) {}
    const char *what() const noexcept override { return msg.c_str(); }
    std::string msg;
};

// ---------------------------------------------------------------- small value types
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
    Point_ operator+(const Point_ &o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_ &o) const { return Point_(x - o.x, y - o.y); }
    Point_ &operator+=(const Point_ &o) { x += o.x; y += o.y; return *this; }
    Point_ &operator-=(const Point_ &o) { x -= o.x; y -= o.y; return *this; }
    Point_ operator*(T s) const { return Point_(x * s, y * s); }
    bool operator==(const Point_ &o) const { return x == o.x && y == o.y; }
};
template <> template <> inline Point_<int>::Point_(const Point_<float> &o) : x(cvRound(o.x)), y(cvRound(o.y)) {}
template <typename T> static inline Point_<T> operator*(T s, const Point_<T> &p) { return Point_<T>(p.x * s, p.y * s); }
template <typename T> static inline Point_<T> operator/(const Point_<T> &p, T s) { return Point_<T>(p.x / s, p.y / s); }
template <typename T> static inline double norm(const Point_<T> &p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    bool operator==(const Size_ &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_ &o) const { return !(*this == o); }
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T> &p, const Size_<T> &s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    Rect_(const Point_<T> &a, const Point_<T> &b) {
        x = std::min(a.x, b.x); y = std::min(a.y, b.y);
        width = std::max(a.x, b.x) - x; height = std::max(a.y, b.y) - y;
    }
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    bool contains(const Point_<T> &p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
    bool operator==(const Rect_ &o) const { return x == o.x && y == o.y && width == o.width && height == o.height; }
    bool operator!=(const Rect_ &o) const { return !(*this == o); }
};
template <typename T> static inline Rect_<T> operator&(const Rect_<T> &a, const Rect_<T> &b) {
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    T x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    if (x2 <= x1 || y2 <= y1) return Rect_<T>();
    return Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}
template <typename T> static inline Rect_<T> &operator&=(Rect_<T> &a, const Rect_<T> &b) { a = a & b; return a; }
template <typename T> static inline Rect_<T> operator|(const Rect_<T> &a, const Rect_<T> &b) {
    if (a.empty()) return b;
    if (b.empty()) return a;
    T x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
    T x2 = std::max(a.x + a.width, b.x + b.width), y2 = std::max(a.y + a.height, b.y + b.height);
    return Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}
typedef Rect_<int> Rect;
typedef Rect_<float> Rect2f;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b) { static_assert(N >= 2, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N >= 3, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(N >= 4, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    Vec operator+(const Vec &o) const { Vec r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(val[i] + o.val[i]); return r; }
    Vec operator-(const Vec &o) const { Vec r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(val[i] - o.val[i]); return r; }
    Vec &operator+=(const Vec &o) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] + o.val[i]); return *this; }
    Vec &operator-=(const Vec &o) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] - o.val[i]); return *this; }
    Vec &operator*=(double s) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] * s); return *this; }
    Vec &operator*=(float s) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] * s); return *this; }
    Vec &operator*=(int s) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] * s); return *this; }
    Vec &operator/=(float s) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>(val[i] / s); return *this; }
    bool operator==(const Vec &o) const { for (int i = 0; i < N; i++) if (val[i] != o.val[i]) return false; return true; }
    bool operator!=(const Vec &o) const { return !(*this == o); }
};
// OpenCV semantics: Vec<T,N> * scalar computes in the scalar's type and saturate_casts back to T.
template <typename T, int N> static inline Vec<T, N> operator*(const Vec<T, N> &v, float s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] * s); return r; }
template <typename T, int N> static inline Vec<T, N> operator*(float s, const Vec<T, N> &v) { return v * s; }
template <typename T, int N> static inline Vec<T, N> operator*(const Vec<T, N> &v, double s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] * s); return r; }
template <typename T, int N> static inline Vec<T, N> operator*(double s, const Vec<T, N> &v) { return v * s; }
template <typename T, int N> static inline Vec<T, N> operator*(const Vec<T, N> &v, int s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] * s); return r; }
template <typename T, int N> static inline Vec<T, N> operator*(int s, const Vec<T, N> &v) { return v * s; }
template <typename T, int N> static inline Vec<T, N> operator/(const Vec<T, N> &v, float s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] / s); return r; }
template <typename T, int N> static inline Vec<T, N> operator/(const Vec<T, N> &v, double s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] / s); return r; }
template <typename T, int N> static inline Vec<T, N> operator/(const Vec<T, N> &v, int s) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(v.val[i] / s); return r; }
template <typename T, int N> static inline Vec<T, N> operator-(const Vec<T, N> &v) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(-v.val[i]); return r; }
template <typename T, int N> static inline double norm(const Vec<T, N> &v) { double s = 0; for (int i = 0; i < N; i++) s += (double)v.val[i] * v.val[i]; return std::sqrt(s); }
typedef Vec<uchar, 3> Vec3b;
typedef Vec<uchar, 4> Vec4b;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<double, 3> Vec3d;

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double a) { val[0] = a; val[1] = val[2] = val[3] = 0; }
    Scalar(double a, double b, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double &operator[](int i) { return val[i]; }
    const double &operator[](int i) const { return val[i]; }
};

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type = 0, maxCount = 0;
    double epsilon = 0;
    TermCriteria() {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> static inline Ptr<T> makePtr(A &&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };

// ---------------------------------------------------------------- Mat
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
                   BORDER_TRANSPARENT = 5, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4,
                          WARP_INVERSE_MAP = 16 };
enum ColorConversionCodes { COLOR_BGR2BGRA = 0, COLOR_BGRA2BGR = 1, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_BGR2HSV = 40,
                            COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56, COLOR_BGR2YCrCb = 36, COLOR_YCrCb2BGR = 38 };
enum ThresholdTypes { THRESH_BINARY = 0, THRESH_BINARY_INV = 1 };
enum RetrievalModes { RETR_EXTERNAL = 0, RETR_LIST = 1 };
enum ContourApproximationModes { CHAIN_APPROX_NONE = 1, CHAIN_APPROX_SIMPLE = 2 };
enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum MorphShapes { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { RANSAC = 8, LMEDS = 4 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };

class Mat;
template <typename T> class Mat_;
struct MatExprZeros;

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    struct Step {
        size_t v = 0;
        operator size_t() const { return v; }
        size_t operator[](int i) const { return i == 0 ? v : 0; }
    } step;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar &s) { create(sz.height, sz.width, type); setTo(s); }
    // wraps user memory (not owned)
    Mat(int r, int c, int type, void *ptr, size_t stp = 0) : rows(r), cols(c), data((uchar *)ptr), type_(type) {
        step.v = stp ? stp : (size_t)c * elemSize();
    }
    Mat(const Mat &m, const Rect &roi) : rows(roi.height), cols(roi.width), type_(m.type_), buf_(m.buf_) {
        if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows)
            throw Exception("mini_cv: ROI outside the matrix");
        step.v = m.step.v;
        data = m.data + (size_t)roi.y * m.step.v + (size_t)roi.x * m.elemSize();
    }
    template <typename T> explicit Mat(const std::vector<T> &v);   // not needed by the reference's live code

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_ && isContinuous()) return;
        rows = r; cols = c; type_ = type;
        step.v = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step.v + 64], std::default_delete<uchar[]>());   // uninitialised, like cv::Mat::create
        data = buf_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { rows = cols = 0; data = nullptr; buf_.reset(); step.v = 0; }

    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)cvDepthSize(depth()) * channels(); }
    size_t elemSize1() const { return (size_t)cvDepthSize(depth()); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step.v == (size_t)cols * elemSize() || rows <= 1; }
    int dims_() const { return 2; }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat &dst) const {
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        if (!(dst.rows == rows && dst.cols == cols && dst.type_ == type_ && dst.data)) dst.create(rows, cols, type_);
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + (size_t)y * dst.step.v, data + (size_t)y * step.v, rb);
    }
    // Mat::copyTo(OutputArray) on an rvalue ROI, e.g. src(r1).copyTo(dst(r2))
    void copyTo(Mat &&dst) const { Mat &d = dst; copyTo(d); }
    void copyTo(Mat &dst, const Mat &mask) const {
        if (!(dst.rows == rows && dst.cols == cols && dst.type_ == type_ && dst.data)) { dst.create(rows, cols, type_); dst.setTo(Scalar()); }
        size_t es = elemSize();
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                if (mask.data[(size_t)y * mask.step.v + x]) std::memcpy(dst.data + (size_t)y * dst.step.v + x * es, data + (size_t)y * step.v + x * es, es);
    }
    void convertTo(Mat &dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        int dd = rtype < 0 ? depth() : (rtype & 7);
        Mat out(rows, cols, CV_MAKETYPE(dd, channels()));
        int n = cols * channels();
        for (int y = 0; y < rows; y++)
            for (int i = 0; i < n; i++) out.setElem(y, i, getElem(y, i) * alpha + beta);
        dst = out;
    }
    Mat &setTo(const Scalar &s) {
        int cn = channels();
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                for (int c = 0; c < cn; c++) setElem(y, x * cn + c, s.val[c]);
        return *this;
    }
    Mat &operator=(const Scalar &s) { return setTo(s); }
    Mat operator()(const Rect &roi) const { return Mat(*this, roi); }
    Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
    Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }

    template <typename T> T &at(int y, int x) { return *(T *)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *(const T *)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
    template <typename T> T &at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T &at(Point p) const { return at<T>(p.y, p.x); }
    template <typename T> T &at(int i) {
        if (rows == 1) return at<T>(0, i);
        if (cols == 1) return at<T>(i, 0);
        return at<T>(i / cols, i % cols);
    }
    template <typename T> const T &at(int i) const { return const_cast<Mat *>(this)->at<T>(i); }
    template <typename T> T *ptr(int y = 0) { return (T *)(data + (size_t)y * step.v); }
    template <typename T> const T *ptr(int y = 0) const { return (const T *)(data + (size_t)y * step.v); }
    uchar *ptr(int y = 0) { return data + (size_t)y * step.v; }
    const uchar *ptr(int y = 0) const { return data + (size_t)y * step.v; }

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, (size_t)r * m.step.v); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(1)); return m; }
    static Mat ones(Size s, int type) { return ones(s.height, s.width, type); }
    static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.setElem(i, i * m.channels(), 1.0); return m; }

    double getElem(int y, int i) const {
        const uchar *p = data + (size_t)y * step.v;
        switch (depth()) {
        case CV_8U: return p[i];
        case CV_8S: return ((const signed char *)p)[i];
        case CV_16U: return ((const ushort *)p)[i];
        case CV_16S: return ((const short *)p)[i];
        case CV_32S: return ((const int *)p)[i];
        case CV_32F: return ((const float *)p)[i];
        default: return ((const double *)p)[i];
        }
    }
    void setElem(int y, int i, double v) {
        uchar *p = data + (size_t)y * step.v;
        switch (depth()) {
        case CV_8U: p[i] = saturate_cast<uchar>(v); break;
        case CV_8S: ((signed char *)p)[i] = (signed char)cvRound(v); break;
        case CV_16U: ((ushort *)p)[i] = (ushort)cvRound(v); break;
        case CV_16S: ((short *)p)[i] = (short)cvRound(v); break;
        case CV_32S: ((int *)p)[i] = cvRound(v); break;
        case CV_32F: ((float *)p)[i] = (float)v; break;
        default: ((double *)p)[i] = v; break;
        }
    }
    mini_cv_mat view() const { mini_cv_mat m; m.data = data; m.rows = rows; m.cols = cols; m.type = type_; m.step = step.v; return m; }

protected:
    int type_ = 0;
    std::shared_ptr<uchar> buf_;
};

// Mat_<T> with the `(Mat_<float>(2,2) << a, b, c, d)` comma initialiser
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat &m) : Mat(m) {}
    T &operator()(int y, int x) { return this->template at<T>(y, x); }
    const T &operator()(int y, int x) const { return this->template at<T>(y, x); }
};
template <typename T> struct MatCommaInitializer_ {
    Mat_<T> m;
    int idx = 0;
    explicit MatCommaInitializer_(const Mat_<T> &mm) : m(mm) {}
    template <typename U> MatCommaInitializer_ &operator,(U v) { m.template at<T>(idx / m.cols, idx % m.cols) = (T)v; idx++; return *this; }
    operator Mat() const { return m; }
    operator Mat_<T>() const { return m; }
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, U v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

static inline Mat operator*(const Mat &m, double s) { Mat r; m.convertTo(r, -1, s, 0.0); return r; }
static inline Mat operator*(double s, const Mat &m) { return m * s; }

// _InputArray / _OutputArray are collapsed onto Mat; noArray() is an empty Mat
typedef const Mat &InputArray;
typedef Mat &OutputArray;
typedef Mat &InputOutputArray;
inline Mat &noArray() { static thread_local Mat none; none.release(); return none; }

inline void mini_cv_check(int rc, const char *what) {
    if (rc != 0) throw Exception(std::string("mini_cv op failed: ") + what);
}
inline const mini_cv_ops *ops() {
    const mini_cv_ops *o = mini_cv_get_ops();
    if (!o) throw Exception("mini_cv: the OpenCV callback table has not been registered");
    return o;
}

// ---------------------------------------------------------------- core / imgproc calls (forwarded to real OpenCV)
inline void setUseOptimized(bool) {}
inline void setNumThreads(int) {}
inline int getNumThreads() { return 1; }

inline void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR) {
    if (src.empty()) throw Exception("resize: empty source");
    if (dsize.width <= 0 || dsize.height <= 0) dsize = Size(cvRound(src.cols * fx), cvRound(src.rows * fy));
    Mat out(dsize.height, dsize.width, src.type());
    mini_cv_mat s = src.view(), d = out.view();
    mini_cv_check(ops()->resize(&s, &d, interp), "resize");
    dst = out;
}
inline void cvtColor(const Mat &src, Mat &dst, int code, int = 0) {
    if (src.empty()) throw Exception("cvtColor: empty source");
    int cn = (code == COLOR_BGR2GRAY) ? 1 : (code == COLOR_BGR2BGRA ? 4 : 3);
    Mat out(src.rows, src.cols, CV_MAKETYPE(src.depth(), cn));
    mini_cv_mat s = src.view(), d = out.view();
    mini_cv_check(ops()->cvt_color(&s, &d, code), "cvtColor");
    dst = out;
}
inline void goodFeaturesToTrack(const Mat &image, std::vector<Point2f> &corners, int maxCorners, double qualityLevel,
                                double minDistance, const Mat &mask = Mat(), int blockSize = 3, bool useHarris = false, double k = 0.04) {
    if (image.empty()) throw Exception("goodFeaturesToTrack: empty image");
    mini_cv_mat s = image.view(), m = mask.view();
    int cap = 1 << 16, n = 0;
    std::vector<float> xy((size_t)cap * 2);
    mini_cv_check(ops()->gftt(&s, mask.empty() ? nullptr : &m, maxCorners, qualityLevel, minDistance, blockSize, useHarris ? 1 : 0, k, xy.data(), cap, &n), "goodFeaturesToTrack");
    corners.resize(n);
    for (int i = 0; i < n; i++) corners[i] = Point2f(xy[2 * i], xy[2 * i + 1]);
}
inline void calcOpticalFlowPyrLK(const Mat &prev, const Mat &next, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                                 std::vector<uchar> &status, std::vector<float> &err, Size win = Size(21, 21), int maxLevel = 3,
                                 TermCriteria crit = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                 double minEigThreshold = 1e-4) {
    if (prev.empty() || next.empty()) throw Exception("calcOpticalFlowPyrLK: empty image");
    if (prev.size() != next.size()) throw Exception("calcOpticalFlowPyrLK: size mismatch");
    int n = (int)prevPts.size();
    nextPts.resize(n); status.resize(n); err.resize(n);
    if (n == 0) return;   // cv: empty outputs for empty input
    mini_cv_mat a = prev.view(), b = next.view();
    mini_cv_check(ops()->pyr_lk(&a, &b, (const float *)prevPts.data(), n, (float *)nextPts.data(), status.data(), err.data(), win.width, win.height,
                                maxLevel, crit.type, crit.maxCount, crit.epsilon, flags, minEigThreshold), "calcOpticalFlowPyrLK");
}
inline Mat estimateAffinePartial2D(const std::vector<Point2f> &from, const std::vector<Point2f> &to, Mat &inliers = noArray(),
                                   int method = RANSAC, double ransacThresh = 3, size_t maxIters = 2000, double confidence = 0.99,
                                   size_t refineIters = 10) {
    int n = (int)from.size();
    if (n != (int)to.size()) throw Exception("estimateAffinePartial2D: size mismatch");
    double M[6];
    std::vector<uchar> mask(std::max(n, 1));
    int ok = 0;
    mini_cv_check(ops()->estimate_affine_partial(n ? (const float *)from.data() : nullptr, n ? (const float *)to.data() : nullptr, n, method, ransacThresh,
                                                 (int)maxIters, confidence, (int)refineIters, M, mask.data(), &ok), "estimateAffinePartial2D");
    if (!ok) return Mat();
    Mat H(2, 3, CV_64FC1);
    for (int i = 0; i < 6; i++) H.at<double>(i / 3, i % 3) = M[i];
    return H;
}
inline void warpAffine(const Mat &src, Mat &dst, const Mat &M, Size dsize, int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT,
                       const Scalar &borderValue = Scalar()) {
    if (src.empty()) throw Exception("warpAffine: empty source");
    if (M.rows != 2 || M.cols != 3) throw Exception("warpAffine: M must be 2x3");
    double m[6];
    for (int i = 0; i < 6; i++) m[i] = M.getElem(i / 3, i % 3);
    Mat out(dsize.height, dsize.width, src.type());
    mini_cv_mat s = src.view(), d = out.view();
    mini_cv_check(ops()->warp_affine(&s, &d, m, M.depth() == CV_32F ? 1 : 0, flags, borderMode, borderValue.val), "warpAffine");
    dst = out;
}
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType, const Scalar &value = Scalar()) {
    if (src.empty()) throw Exception("copyMakeBorder: empty source");
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    mini_cv_mat s = src.view(), d = out.view();
    mini_cv_check(ops()->copy_make_border(&s, &d, top, bottom, left, right, borderType, value.val), "copyMakeBorder");
    dst = out;
}
inline void addWeighted(const Mat &a, double alpha, const Mat &b, double beta, double gamma, Mat &dst, int = -1) {
    if (a.size() != b.size() || a.type() != b.type()) throw Exception("addWeighted: size/type mismatch");
    Mat out(a.rows, a.cols, a.type());
    mini_cv_mat x = a.view(), y = b.view(), d = out.view();
    mini_cv_check(ops()->add_weighted(&x, alpha, &y, beta, gamma, &d), "addWeighted");
    dst = out;
}
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type) {
    Mat out(src.rows, src.cols, src.type());
    mini_cv_mat s = src.view(), d = out.view();
    mini_cv_check(ops()->threshold(&s, &d, thresh, maxval, type), "threshold");
    dst = out;
    return thresh;
}
inline void findContours(const Mat &image, std::vector<std::vector<Point>> &contours, int mode, int method) {
    mini_cv_mat s = image.view();
    int total_cap = 1 << 20, cont_cap = 1 << 16, ncont = 0;
    std::vector<int> pts((size_t)total_cap * 2), lens(cont_cap);
    mini_cv_check(ops()->find_contours(&s, mode, method, pts.data(), total_cap, lens.data(), cont_cap, &ncont), "findContours");
    contours.clear();
    size_t k = 0;
    for (int c = 0; c < ncont; c++) {
        std::vector<Point> v(lens[c]);
        for (int i = 0; i < lens[c]; i++, k++) v[i] = Point(pts[2 * k], pts[2 * k + 1]);
        contours.push_back(std::move(v));
    }
}
inline double contourArea(const std::vector<Point> &c) {   // shoelace, as cv::contourArea(oriented=false)
    double a = 0;
    int n = (int)c.size();
    if (n == 0) return 0;
    Point2f prev((float)c[n - 1].x, (float)c[n - 1].y);
    for (int i = 0; i < n; i++) {
        Point2f p((float)c[i].x, (float)c[i].y);
        a += (double)prev.x * p.y - (double)prev.y * p.x;
        prev = p;
    }
    return std::fabs(a * 0.5);
}
inline Rect boundingRect(const std::vector<Point> &pts) {
    if (pts.empty()) return Rect();
    int x0 = pts[0].x, x1 = pts[0].x, y0 = pts[0].y, y1 = pts[0].y;
    for (auto &p : pts) { x0 = std::min(x0, p.x); x1 = std::max(x1, p.x); y0 = std::min(y0, p.y); y1 = std::max(y1, p.y); }
    return Rect(x0, y0, x1 - x0 + 1, y1 - y0 + 1);
}
// drawing: only used by debug overlays that the parity paths never reach
inline void line(Mat &, Point, Point, const Scalar &, int = 1, int = LINE_8, int = 0) { throw Exception("mini_cv: cv::line is not provided"); }
inline void circle(Mat &, Point, int, const Scalar &, int = 1, int = LINE_8, int = 0) { throw Exception("mini_cv: cv::circle is not provided"); }
inline void rectangle(Mat &img, Rect r, const Scalar &color, int thickness = 1, int = LINE_8, int = 0) {
    if (thickness >= 0) throw Exception("mini_cv: only filled cv::rectangle is provided");
    Rect c = r & Rect(0, 0, img.cols, img.rows);
    if (c.empty()) return;
    Mat roi = img(c);
    roi.setTo(color);
}
inline void rectangle(Mat &img, Point a, Point b, const Scalar &color, int thickness = 1, int lt = LINE_8, int shift = 0) {
    // cv::rectangle(pt1, pt2) includes both corners
    rectangle(img, Rect(std::min(a.x, b.x), std::min(a.y, b.y), std::abs(a.x - b.x) + 1, std::abs(a.y - b.y) + 1), color, thickness, lt, shift);
}

// ---------------------------------------------------------------- classes the reference instantiates
class CLAHE {
public:
    virtual ~CLAHE() {}
    virtual void apply(const Mat &, Mat &) { throw Exception("mini_cv: CLAHE is not provided (dead code in the reference)"); }
    virtual void setClipLimit(double) {}
    virtual void setTilesGridSize(Size) {}
};
inline Ptr<CLAHE> createCLAHE(double = 40.0, Size = Size(8, 8)) { return std::make_shared<CLAHE>(); }

class Feature2D {
public:
    virtual ~Feature2D() {}
    virtual void detect(const Mat &, std::vector<KeyPoint> &, const Mat & = Mat()) { throw Exception("mini_cv: Feature2D::detect is not provided (dead code in the reference)"); }
};
class ORB : public Feature2D {
public:
    static Ptr<ORB> create(int = 500, float = 1.2f, int = 8, int = 31, int = 0, int = 2, int = 0, int = 31, int = 20) { return std::make_shared<ORB>(); }
};
class BRISK : public Feature2D {
public:
    static Ptr<BRISK> create(int = 30, int = 3, float = 1.0f) { return std::make_shared<BRISK>(); }
};
inline void FAST(const Mat &, std::vector<KeyPoint> &, int, bool = true) { throw Exception("mini_cv: FAST is not provided (dead code in the reference)"); }

// cv::KalmanFilter: the state lives here, predict()/correct() are done by cv2.KalmanFilter through the callback
class KalmanFilter {
public:
    KalmanFilter() {}
    KalmanFilter(int dynamParams, int measureParams, int controlParams = 0, int type = CV_32F) { init(dynamParams, measureParams, controlParams, type); }
    void init(int DP, int MP, int CP = 0, int type = CV_32F) {
        if (type != CV_32F) throw Exception("mini_cv: KalmanFilter supports CV_32F only");
        dp = DP; mp = MP; cp = CP;
        statePre = Mat::zeros(DP, 1, type); statePost = Mat::zeros(DP, 1, type);
        transitionMatrix = Mat::eye(DP, DP, type);
        processNoiseCov = Mat::eye(DP, DP, type);
        measurementMatrix = Mat::zeros(MP, DP, type);
        measurementNoiseCov = Mat::eye(MP, MP, type);
        errorCovPre = Mat::zeros(DP, DP, type); errorCovPost = Mat::zeros(DP, DP, type);
        gain = Mat::zeros(DP, MP, type);
        handle = -1;
    }
    ~KalmanFilter() { if (handle >= 0 && mini_cv_get_ops()) mini_cv_get_ops()->kalman_release(handle); }
    const Mat &predict(const Mat & = Mat()) { sync(); mini_cv_check(ops()->kalman_predict(handle, state_views()), "KalmanFilter::predict"); return statePre; }
    const Mat &correct(const Mat &measurement) {
        sync();
        mini_cv_mat m = measurement.view();
        mini_cv_check(ops()->kalman_correct(handle, &m, state_views()), "KalmanFilter::correct");
        return statePost;
    }
    Mat statePre, statePost, transitionMatrix, controlMatrix, measurementMatrix, processNoiseCov, measurementNoiseCov, errorCovPre, gain, errorCovPost;

private:
    // every call pushes the (possibly user-edited) matrices to the cv2 object and reads the state back
    mini_cv_mat *state_views() {
        Mat *all[9] = {&statePre, &statePost, &transitionMatrix, &measurementMatrix, &processNoiseCov, &measurementNoiseCov, &errorCovPre, &gain, &errorCovPost};
        for (int i = 0; i < 9; i++) views_[i] = all[i]->view();
        return views_;
    }
    void sync() {
        if (handle < 0) mini_cv_check(ops()->kalman_create(dp, mp, cp, &handle), "KalmanFilter::create");
    }
    int dp = 0, mp = 0, cp = 0, handle = -1;
    mini_cv_mat views_[9];
};

}  // namespace cv

#endif
