// Minimal OpenCV stand-in for the ORACLE build only (test infrastructure, never shipped).
//
// The reference (Asherchi/Stereo_Match_Traditional) links OpenCV 3.1.0
// (opencv_world310.lib, e.g. SAD/SAD.vcxproj:127) which is not available offline.
// This header implements just the subset of cv:: that the reference's hot-path
// headers touch, with OpenCV's documented semantics, so that the UNMODIFIED
// reference sources compile under g++ and can serve as the parity oracle
// (oracle/ref_harness.cpp).  Display / file-IO entry points are declared but abort.
//
// Semantics that matter for parity and where they are used by the reference:
//   * cv::abs(a - b) on 8-bit Mats == absdiff (Sad.h:17)            -> SubProxy below
//   * cv::sum on CV_8U / CV_32S / CV_32F / CV_64F, per channel 0    (Sad.h:18, ASW.h:248-255)
//   * Mat ROI via (Range, Range) keeps the parent's step            (Sad.h:119, NCC.h:78, ASW.h:341)
//   * copyMakeBorder(BORDER_REPLICATE)                              (SADmain.cpp:47, ASWeight.cpp:54)
//   * convertTo(CV_64F), Mat::mul on CV_64F                         (ASW.h:247-254)
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace cv {

typedef unsigned char uchar;
typedef unsigned short ushort;

enum { CV_8U = 0, CV_8S = 1, CV_16U = 2, CV_16S = 3, CV_32S = 4, CV_32F = 5, CV_64F = 6 };
#define SMT_CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
enum {
    CV_8UC1 = SMT_CV_MAKETYPE(CV_8U, 1), CV_8UC3 = SMT_CV_MAKETYPE(CV_8U, 3),
    CV_16UC1 = SMT_CV_MAKETYPE(CV_16U, 1), CV_16SC1 = SMT_CV_MAKETYPE(CV_16S, 1),
    CV_32SC1 = SMT_CV_MAKETYPE(CV_32S, 1),
    CV_32FC1 = SMT_CV_MAKETYPE(CV_32F, 1), CV_32FC3 = SMT_CV_MAKETYPE(CV_32F, 3),
    CV_64FC1 = SMT_CV_MAKETYPE(CV_64F, 1), CV_64FC3 = SMT_CV_MAKETYPE(CV_64F, 3)
};
enum { CV_BGR2GRAY = 6, CV_BGR2Lab = 44, COLOR_BGR2GRAY = 6, COLOR_BGR2Lab = 44 };
enum { CV_WINDOW_NORMAL = 0, WINDOW_NORMAL = 0, CV_WINDOW_AUTOSIZE = 1, WINDOW_AUTOSIZE = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_MINMAX = 32 };
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };

inline int smt_depth_size(int depth) {
    static const int sz[7] = {1, 1, 2, 2, 4, 4, 8};
    return sz[depth & 7];
}

[[noreturn]] inline void smt_unsupported(const char* what) {
    std::fprintf(stderr, "cvshim: '%s' is not implemented in the oracle stand-in\n", what);
    std::abort();
}

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3-vector"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;
typedef Vec<int, 3> Vec3i;

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

class Mat;
struct SubProxy;  // result of (Mat - Mat); only cv::abs() or conversion to Mat may consume it

class Mat {
public:
    int flags;  // == type()
    int rows, cols;
    size_t step;  // bytes per row of the *parent* allocation
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() : flags(0), rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
    // user-allocated data: no ownership (OpenCV semantics)
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)ext) {
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    Mat(const SubProxy& p);
    Mat& operator=(const SubProxy& p);

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type && buf) return;
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, (uchar)0);
        data = buf->data();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); std::memset(m.data, 0, (size_t)r * m.step); return m; }
    static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(1)); return m; }

    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> 3) + 1; }
    size_t elemSize() const { return (size_t)smt_depth_size(depth()) * channels(); }
    size_t elemSize1() const { return (size_t)smt_depth_size(depth()); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    template <typename T> T* ptr(int i = 0) { return reinterpret_cast<T*>(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return reinterpret_cast<const T*>(data + (size_t)i * step); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T& at(int i, int j) { return reinterpret_cast<T*>(data + (size_t)i * step)[j]; }
    template <typename T> const T& at(int i, int j) const { return reinterpret_cast<const T*>(data + (size_t)i * step)[j]; }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }

    // ROI view: shares the buffer, keeps the parent's step
    Mat operator()(Range rr, Range cr) const {
        Mat v;
        int r0 = rr.start == INT32_MIN ? 0 : rr.start, r1 = rr.end == INT32_MAX ? rows : rr.end;
        int c0 = cr.start == INT32_MIN ? 0 : cr.start, c1 = cr.end == INT32_MAX ? cols : cr.end;
        v.flags = flags; v.rows = r1 - r0; v.cols = c1 - c0; v.step = step; v.buf = buf;
        v.data = data + (size_t)r0 * step + (size_t)c0 * elemSize();
        return v;
    }
    Mat operator()(const Rect& r) const { return (*this)(Range(r.y, r.y + r.height), Range(r.x, r.x + r.width)); }
    Mat row(int i) const { return (*this)(Range(i, i + 1), Range::all()); }
    Mat col(int j) const { return (*this)(Range::all(), Range(j, j + 1)); }
    Mat rowRange(int a, int b) const { return (*this)(Range(a, b), Range::all()); }
    Mat colRange(int a, int b) const { return (*this)(Range::all(), Range(a, b)); }

    Mat clone() const {
        Mat m; m.create(rows, cols, flags);
        for (int i = 0; i < rows; ++i) std::memcpy(m.ptr(i), ptr(i), (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }

    double getAsDouble(int i, int j, int ch = 0) const {
        const uchar* p = data + (size_t)i * step + (size_t)j * elemSize() + (size_t)ch * elemSize1();
        switch (depth()) {
            case CV_8U: return *p;
            case CV_8S: return *(const signed char*)p;
            case CV_16U: return *(const ushort*)p;
            case CV_16S: return *(const short*)p;
            case CV_32S: return *(const int*)p;
            case CV_32F: return *(const float*)p;
            default: return *(const double*)p;
        }
    }
    static long long smt_round(double v) { return (long long)std::nearbyint(v); }  // cvRound: ties-to-even
    void setFromDouble(int i, int j, int ch, double v) {
        uchar* p = data + (size_t)i * step + (size_t)j * elemSize() + (size_t)ch * elemSize1();
        switch (depth()) {  // saturate_cast semantics
            case CV_8U: { long long r = smt_round(v); *p = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r); break; }
            case CV_8S: { long long r = smt_round(v); *(signed char*)p = (signed char)(r < -128 ? -128 : r > 127 ? 127 : r); break; }
            case CV_16U: { long long r = smt_round(v); *(ushort*)p = (ushort)(r < 0 ? 0 : r > 65535 ? 65535 : r); break; }
            case CV_16S: { long long r = smt_round(v); *(short*)p = (short)(r < -32768 ? -32768 : r > 32767 ? 32767 : r); break; }
            case CV_32S: *(int*)p = (int)smt_round(v); break;
            case CV_32F: *(float*)p = (float)v; break;
            default: *(double*)p = v; break;
        }
    }
    Mat& setTo(const Scalar& s) {
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) for (int c = 0; c < channels(); ++c) setFromDouble(i, j, c, s.val[c < 4 ? c : 3]);
        return *this;
    }
    Mat& operator=(const Scalar& s) { return setTo(s); }

    // dst may alias *this (the reference does error.convertTo(error, CV_64F), ASW.h:253)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        int cn = channels();
        int ddepth = rtype < 0 ? depth() : (rtype & 7);
        Mat out(rows, cols, SMT_CV_MAKETYPE(ddepth, cn));
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) for (int c = 0; c < cn; ++c)
            out.setFromDouble(i, j, c, getAsDouble(i, j, c) * alpha + beta);
        dst = out;
    }
    // element-wise product, result has the operands' type
    Mat mul(const Mat& o, double scale = 1.0) const {
        assert(rows == o.rows && cols == o.cols && flags == o.flags);
        Mat out(rows, cols, flags);
        int cn = channels();
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) for (int c = 0; c < cn; ++c)
            out.setFromDouble(i, j, c, getAsDouble(i, j, c) * o.getAsDouble(i, j, c) * scale);
        return out;
    }
    Mat reshape(int, int = 0) const { smt_unsupported("Mat::reshape"); }
};

struct SubProxy { const Mat& a; const Mat& b; };
inline SubProxy operator-(const Mat& a, const Mat& b) { return SubProxy{a, b}; }

inline void absdiff(const Mat& a, const Mat& b, Mat& dst) {
    assert(a.rows == b.rows && a.cols == b.cols && a.flags == b.flags);
    Mat out(a.rows, a.cols, a.flags);
    int cn = a.channels();
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) for (int c = 0; c < cn; ++c)
        out.setFromDouble(i, j, c, std::fabs(a.getAsDouble(i, j, c) - b.getAsDouble(i, j, c)));
    dst = out;
}
// OpenCV: abs(A - B) is evaluated as absdiff(A, B) (MatExpr special case) -- no saturation of the difference.
inline Mat abs(const SubProxy& p) { Mat d; absdiff(p.a, p.b, d); return d; }
// plain saturating subtraction if someone materialises the proxy
inline Mat smt_subtract(const Mat& a, const Mat& b) {
    Mat out(a.rows, a.cols, a.flags);
    int cn = a.channels();
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) for (int c = 0; c < cn; ++c)
        out.setFromDouble(i, j, c, a.getAsDouble(i, j, c) - b.getAsDouble(i, j, c));
    return out;
}
inline Mat::Mat(const SubProxy& p) : Mat() { *this = smt_subtract(p.a, p.b); }
inline Mat& Mat::operator=(const SubProxy& p) { *this = smt_subtract(p.a, p.b); return *this; }
inline Mat abs(const Mat& m) {
    Mat out(m.rows, m.cols, m.flags);
    int cn = m.channels();
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) for (int c = 0; c < cn; ++c)
        out.setFromDouble(i, j, c, std::fabs(m.getAsDouble(i, j, c)));
    return out;
}

inline Scalar sum(const Mat& m) {
    Scalar s;
    int cn = m.channels();
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) for (int c = 0; c < cn && c < 4; ++c)
        s.val[c] += m.getAsDouble(i, j, c);
    return s;
}
inline Scalar mean(const Mat& m) {
    Scalar s = sum(m);
    double n = (double)m.rows * m.cols;
    for (int c = 0; c < 4; ++c) s.val[c] = n > 0 ? s.val[c] / n : 0;
    return s;
}

inline Mat operator/(const Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0 / d, 0); return o; }
inline Mat operator*(const Mat& m, double d) { Mat o; m.convertTo(o, -1, d, 0); return o; }
inline Mat operator*(double d, const Mat& m) { return m * d; }
inline Mat operator+(const Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0, d); return o; }
inline Mat operator-(const Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0, -d); return o; }
inline Mat& operator+=(Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0, d); m = o; return m; }
inline Mat& operator-=(Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0, -d); m = o; return m; }
inline Mat& operator/=(Mat& m, double d) { Mat o; m.convertTo(o, -1, 1.0 / d, 0); m = o; return m; }
inline Mat& operator*=(Mat& m, double d) { Mat o; m.convertTo(o, -1, d, 0); m = o; return m; }
inline Mat operator+(const Mat& a, const Mat& b) {
    Mat out(a.rows, a.cols, a.flags);
    int cn = a.channels();
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) for (int c = 0; c < cn; ++c)
        out.setFromDouble(i, j, c, a.getAsDouble(i, j, c) + b.getAsDouble(i, j, c));
    return out;
}
inline Mat& operator+=(Mat& a, const Mat& b) { a = a + b; return a; }

inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                           const Scalar& = Scalar()) {
    if (borderType != BORDER_REPLICATE) smt_unsupported("copyMakeBorder(non-replicate)");
    Mat out(src.rows + top + bottom, src.cols + left + right, src.flags);
    size_t es = src.elemSize();
    for (int i = 0; i < out.rows; ++i) {
        int si = std::min(std::max(i - top, 0), src.rows - 1);
        for (int j = 0; j < out.cols; ++j) {
            int sj = std::min(std::max(j - left, 0), src.cols - 1);
            std::memcpy(out.ptr(i) + (size_t)j * es, src.ptr(si) + (size_t)sj * es, es);
        }
    }
    dst = out;  // src may alias dst (CBLSM.cpp:125)
}

// BGR -> gray with OpenCV's 8-bit fixed-point coefficients (B 1868, G 9617, R 4899, >>14 with rounding).
inline void smt_bgr2gray(const Mat& src, Mat& dst) {
    if (src.type() != CV_8UC3) smt_unsupported("cvtColor(BGR2GRAY) on non-8UC3");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int i = 0; i < src.rows; ++i) for (int j = 0; j < src.cols; ++j) {
        const Vec3b& p = src.at<Vec3b>(i, j);
        out.at<uchar>(i, j) = (uchar)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + (1 << 13)) >> 14);
    }
    dst = out;
}
inline void cvtColor(const Mat& src, Mat& dst, int code, int = 0) {
    if (code == CV_BGR2GRAY) { smt_bgr2gray(src, dst); return; }
    smt_unsupported("cvtColor(code other than BGR2GRAY)");
}
// The reference passes a const Mat& as destination (Sad.h:103, CBLSM.h:415); OpenCV's OutputArray accepts that.
inline void cvtColor(const Mat& src, const Mat& dst, int code, int dcn = 0) { cvtColor(src, const_cast<Mat&>(dst), code, dcn); }

// ---- display / IO / filters only used by the reference's out-of-scope display stages: declared, abort if called
inline void medianBlur(const Mat&, Mat&, int) { smt_unsupported("medianBlur"); }
inline void normalize(const Mat&, Mat&, double = 1, double = 0, int = NORM_L2, int = -1) { smt_unsupported("normalize"); }
inline void normalize(const Mat&, const Mat&, double = 1, double = 0, int = NORM_L2, int = -1) { smt_unsupported("normalize"); }
inline void filterSpeckles(Mat&, double, int, double) { smt_unsupported("filterSpeckles"); }
inline Mat imread(const std::string&, int = 1) { smt_unsupported("imread"); }
inline bool imwrite(const std::string&, const Mat&) { smt_unsupported("imwrite"); }
inline void imshow(const std::string&, const Mat&) { smt_unsupported("imshow"); }
inline void namedWindow(const std::string&, int = 1) { smt_unsupported("namedWindow"); }
inline int waitKey(int = 0) { smt_unsupported("waitKey"); }
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = 1) { smt_unsupported("resize"); }
inline void minMaxLoc(const Mat&, double*, double* = nullptr, Point* = nullptr, Point* = nullptr) { smt_unsupported("minMaxLoc"); }
inline void split(const Mat&, std::vector<Mat>&) { smt_unsupported("split"); }
inline void merge(const std::vector<Mat>&, Mat&) { smt_unsupported("merge"); }

}  // namespace cv

// OpenCV's C-API era headers leak these into the global namespace; the reference relies on it (CV_8UC1, uchar ...).
using cv::uchar;
using cv::ushort;
using cv::CV_8U; using cv::CV_8S; using cv::CV_16U; using cv::CV_16S; using cv::CV_32S; using cv::CV_32F; using cv::CV_64F;
using cv::CV_8UC1; using cv::CV_8UC3; using cv::CV_16UC1; using cv::CV_16SC1; using cv::CV_32SC1;
using cv::CV_32FC1; using cv::CV_32FC3; using cv::CV_64FC1; using cv::CV_64FC3;
using cv::CV_BGR2GRAY; using cv::CV_BGR2Lab; using cv::CV_WINDOW_NORMAL; using cv::CV_WINDOW_AUTOSIZE;
