// ORACLE (test infrastructure, not product code).
// cvshim.h — a minimal `cv::` namespace, just large enough for the UNMODIFIED reference sources on the hot path
// (src/ORBExtractor.cpp, src/CamModelGeneral.cpp, src/ORBMatcher.cpp, src/Frame.cpp, ThirdParty/DBoW2) to compile in a
// container that has no OpenCV C++ headers. Every arithmetic primitive is backed by oracle/cvprim.h, whose models are pinned
// bit-exact against opencv-python 4.13 in tests/test_oracle_cv2.py (remap, resize, FAST-9/16 + NMS, 7x7 Gaussian, fastAtan2,
// cvRound, and the float gemm used by `R*x+t`). Nothing here is a copy of OpenCV code; it is an API-shaped wrapper around our
// own primitive models. The build recipe is oracle/Makefile (target _ref/libref.so).
#ifndef ORACLE_CVSHIM_H
#define ORACLE_CVSHIM_H
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
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <fstream>
#include <list>
#include <string>
#include <vector>
// The shim (like the OpenCV library it stands in for) is never compiled with FMA contraction, whatever flags the including
// reference translation unit uses (oracle/Makefile builds a -ffp-contract=fast variant to quantify contraction in the reference code).
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#include "../cvprim.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_PI 3.1415926535897932384626433832795
#define CV_GRAY2BGR 8
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_Assert(x) assert(x)

namespace cv {

static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

template <class T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };
static inline size_t elem_size(int type) { return type == CV_8U ? 1 : (type == CV_64F ? 8 : 4); }

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
template <class T> static inline Point_<T> operator*(const Point_<T>& p, float s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
template <class T> static inline Point_<T> operator*(const Point_<T>& p, double s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
template <class T> static inline Point_<T> operator*(const Point_<T>& p, int s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
template <class T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> static inline std::ostream& operator<<(std::ostream& o, const Point_<T>& p) { return o << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point2i; typedef Point2i Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

class Mat;
template <class T, int n> struct Vec {
    T val[n];
    Vec() { for (int i = 0; i < n; i++) val[i] = 0; }
    Vec(T a, T b) { static_assert(n == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(n == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    explicit Vec(const Mat& m);
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T dot(const Vec& o) const { T s = 0; for (int i = 0; i < n; i++) s += val[i] * o.val[i]; return s; }   // saturate_cast<T>(sum) for float: plain fp32 accumulation
    Vec cross(const Vec& o) const { static_assert(n == 3, ""); return Vec(val[1] * o.val[2] - val[2] * o.val[1], val[2] * o.val[0] - val[0] * o.val[2], val[0] * o.val[1] - val[1] * o.val[0]); }
};
typedef Vec<float, 2> Vec2f; typedef Vec<double, 2> Vec2d; typedef Vec<float, 3> Vec3f; typedef Vec<double, 3> Vec3d;
template <class T, int n> static inline Vec<T, n> operator-(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = a.val[i] - b.val[i]; return r; }
template <class T, int n> static inline Vec<T, n> operator+(const Vec<T, n>& a, const Vec<T, n>& b) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = a.val[i] + b.val[i]; return r; }
// Vec * scalar (Matx_ScaleOp): saturate_cast<T>(a[i] * alpha) with alpha of the given scalar type (int / float / double overloads)
template <class T, int n> static inline Vec<T, n> operator*(const Vec<T, n>& a, int s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
template <class T, int n> static inline Vec<T, n> operator*(const Vec<T, n>& a, float s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
template <class T, int n> static inline Vec<T, n> operator*(const Vec<T, n>& a, double s) { Vec<T, n> r; for (int i = 0; i < n; i++) r.val[i] = (T)(a.val[i] * s); return r; }
template <class T, int n> static inline Vec<T, n> operator*(float s, const Vec<T, n>& a) { return a * s; }
template <class T, int n> static inline Vec<T, n> operator*(double s, const Vec<T, n>& a) { return a * s; }
template <class T, int n> static inline Vec<T, n> operator/(const Vec<T, n>& a, float s) { return a * (1.f / s); }
template <class T, int n> static inline Vec<T, n> operator/(const Vec<T, n>& a, double s) { return a * (1. / s); }
// cv::norm(Vec<T,n>) : sqrt of the sum of squares accumulated in double (normL2Sqr<T,double>)
template <class T, int n> static inline double norm(const Vec<T, n>& v) { double s = 0; for (int i = 0; i < n; i++) s += (double)v.val[i] * (double)v.val[i]; return std::sqrt(s); }
using std::sqrt; using std::abs; using std::min; using std::max; using std::exp; using std::pow; using std::log; using std::swap;   // like opencv2/core/cvstd.hpp

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class _OutputArray;
struct MatScaleExpr;
// Mat::zeros / ones / eye return an initializer EXPRESSION in OpenCV: assigning it to an existing Mat of the same size and type fills
// that Mat in place (m.create() is a no-op, then m = Scalar), which ORBextractor's computeDescriptors relies on for its row-range view.
struct MatInit { int rows, cols, type, kind; };
// Reference-counted dense 2-D matrix with ROI views (CV_8U / CV_32S / CV_32F / CV_64F, single channel).
class Mat {
public:
    int rows, cols, flags; size_t step; uchar* data;
    Mat() : rows(0), cols(0), flags(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), flags(0), step(0), data(nullptr) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), flags(0), step(0), data(nullptr) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), flags(type), step(s ? s : (size_t)c * elem_size(type)), data((uchar*)d) {}
    Mat(int r, int c, int type, const Scalar& s) : rows(0), cols(0), flags(0), step(0), data(nullptr) { create(r, c, type); setTo(s.val[0]); }
    Mat(const MatInit& e) : rows(0), cols(0), flags(0), step(0), data(nullptr) { *this = e; }
    Mat& operator=(const MatInit& e) {
        create(e.rows, e.cols, e.type);
        setTo(e.kind == 1 ? 1.0 : 0.0);
        if (e.kind == 2) for (int i = 0; i < std::min(e.rows, e.cols); i++) set(i, i, 1);
        return *this;
    }
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type) return;   // cv::Mat::create keeps a matching buffer (also a ROI)
        rows = r; cols = c; flags = type; step = (size_t)c * elem_size(type);
        buf_ = std::shared_ptr<uchar>(new uchar[std::max<size_t>((size_t)r * step, 1)], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return flags; }
    int depth() const { return flags; }
    int channels() const { return 1; }
    size_t elemSize() const { return elem_size(flags); }
    size_t step1() const { return step / elem_size(flags); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elem_size(flags) || rows <= 1; }
    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m; m.rows = r1 - r0; m.cols = c1 - c0; m.flags = flags; m.step = step; m.data = data + (size_t)r0 * step + (size_t)c0 * elem_size(flags); m.buf_ = buf_;
        return m;
    }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat operator()(const Rect& r) const { return view(r.y, r.y + r.height, r.x, r.x + r.width); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elem_size(flags);
        for (int r = 0; r < rows; r++) std::memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rb);
    }
    void copyTo(const _OutputArray& dst) const;   // also binds temporaries such as `a.copyTo(b.rowRange(0,3))`
    void setTo(double v) {
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, v);
    }
    void resize(size_t nrows) {   // cv::Mat::resize(sz): keeps the first rows
        if ((int)nrows == rows) return;
        if ((int)nrows < rows) { rows = (int)nrows; return; }
        Mat m(int(nrows), cols, flags); const size_t rb = (size_t)cols * elem_size(flags);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, rb);
        *this = m;
    }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }   // vector access: row or column vector
    template <class T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    double get(int r, int c) const {
        switch (flags) { case CV_8U: return at<uchar>(r, c); case CV_32S: return at<int>(r, c); case CV_32F: return at<float>(r, c); default: return at<double>(r, c); }
    }
    void set(int r, int c, double v) {
        switch (flags) { case CV_8U: at<uchar>(r, c) = (uchar)v; break; case CV_32S: at<int>(r, c) = (int)v; break; case CV_32F: at<float>(r, c) = (float)v; break; default: at<double>(r, c) = v; }
    }
    static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0}; }
    static MatInit zeros(Size s, int type) { return MatInit{s.height, s.width, type, 0}; }
    static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1}; }
    static MatInit eye(int r, int c, int type) { return MatInit{r, c, type, 2}; }
    struct MatScaleExpr t() const;   // lazy, like cv::MatExpr
    // cv::Mat::dot: element products accumulated in double (dotProd_<float> returns double)
    double dot(const Mat& o) const {
        assert(total() == o.total());
        double s = 0;
        if (rows == o.rows) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += get(r, c) * o.get(r, c); }
        else { for (size_t i = 0; i < total(); i++) s += get((int)(i / cols), (int)(i % cols)) * o.get((int)(i / o.cols), (int)(i % o.cols)); }
        return s;
    }
private:
    std::shared_ptr<uchar> buf_;
};

template <class T, int n> Vec<T, n>::Vec(const Mat& m) { assert((int)m.total() == n); for (int i = 0; i < n; i++) val[i] = (T)(m.rows == 1 ? m.get(0, i) : m.get(i, 0)); }

// ---- matrix arithmetic. OpenCV evaluates `A*B+C` as ONE gemm, D = alpha*op(A)*op(B) + beta*C, and cv::Mat::t() / unary minus / scalar
// factors stay lazy (MatExpr), so `-R.t()*t` is a single gemm with GEMM_1_T and alpha = -1. Pinned against cv2.gemm 4.13
// (tests/test_oracle_cv2.py::test_gemm_models):
//   * flags == 0, CV_32F, 2 <= len <= 4 and len == D.cols or D.rows (the "small matrix" path: R*x+t, R*R): the inner product is accumulated in
//     FLOAT, left to right, without FMA, then D = (float)((double)t*alpha + (double)c*beta);
//   * everything else (any transposed operand, larger sizes, CV_64F): inner product accumulated in double.
struct MatMulExpr {
    Mat a, b; double alpha; bool ta, tb;
    Mat eval(const Mat* c = nullptr, double beta = 0) const {
        const int ar = ta ? a.cols : a.rows, ac = ta ? a.rows : a.cols, br = tb ? b.cols : b.rows, bc = tb ? b.rows : b.cols;
        assert(ac == br && a.type() == b.type());
        (void)br;
        Mat d(ar, bc, a.type());
        const int len = ac;
        const bool smallF = !ta && !tb && a.type() == CV_32F && len >= 2 && len <= 4 && (len == bc || len == ar);
        for (int i = 0; i < ar; i++) for (int j = 0; j < bc; j++) {
            double v;
            if (smallF) {
                float t = a.at<float>(i, 0) * b.at<float>(0, j);
                for (int k = 1; k < len; k++) t = t + a.at<float>(i, k) * b.at<float>(k, j);
                v = (double)t * alpha;
            } else {
                double sacc = 0;
                for (int k = 0; k < len; k++) sacc += (ta ? a.get(k, i) : a.get(i, k)) * (tb ? b.get(j, k) : b.get(k, j));
                v = sacc * alpha;
            }
            if (c) v += c->get(i, j) * beta;
            d.set(i, j, v);
        }
        return d;
    }
    operator Mat() const { return eval(); }
};
// alpha * A or alpha * A^T, unevaluated
struct MatScaleExpr {
    Mat a; double alpha; bool t;
    operator Mat() const {
        Mat d(t ? a.cols : a.rows, t ? a.rows : a.cols, a.type());
        for (int i = 0; i < d.rows; i++) for (int j = 0; j < d.cols; j++) { const double v = t ? a.get(j, i) : a.get(i, j); d.set(i, j, alpha == 1.0 ? v : v * alpha); }
        return d;
    }
    MatScaleExpr t_() const { return MatScaleExpr{a, alpha, !t}; }
};
inline MatScaleExpr Mat::t() const { return MatScaleExpr{*this, 1.0, true}; }
static inline MatMulExpr operator*(const Mat& a, const Mat& b) { return MatMulExpr{a, b, 1.0, false, false}; }
static inline MatMulExpr operator*(const MatScaleExpr& s, const Mat& b) { return MatMulExpr{s.a, b, s.alpha, s.t, false}; }
static inline MatMulExpr operator*(const Mat& a, const MatScaleExpr& s) { return MatMulExpr{a, s.a, s.alpha, false, s.t}; }
static inline MatMulExpr operator*(const MatMulExpr& e, const Mat& b) { return MatMulExpr{e.eval(), b, 1.0, false, false}; }
static inline Mat operator+(const MatMulExpr& e, const Mat& c) { return e.eval(&c, 1.0); }
static inline Mat operator-(const MatMulExpr& e, const Mat& c) { return e.eval(&c, -1.0); }
static inline Mat operator+(const Mat& c, const MatMulExpr& e) { return e.eval(&c, 1.0); }
static inline MatMulExpr operator-(const MatMulExpr& e) { return MatMulExpr{e.a, e.b, -e.alpha, e.ta, e.tb}; }
static inline MatScaleExpr operator-(const Mat& a) { return MatScaleExpr{a, -1.0, false}; }
static inline MatScaleExpr operator-(const MatScaleExpr& s) { return MatScaleExpr{s.a, -s.alpha, s.t}; }
static inline MatScaleExpr operator*(double k, const Mat& a) { return MatScaleExpr{a, k, false}; }
static inline MatScaleExpr operator*(const Mat& a, double k) { return MatScaleExpr{a, k, false}; }
static inline MatScaleExpr operator*(double k, const MatScaleExpr& s) { return MatScaleExpr{s.a, k * s.alpha, s.t}; }
static inline MatScaleExpr operator/(const Mat& a, double k) { return MatScaleExpr{a, 1.0 / k, false}; }
static inline Mat binop(const Mat& a, const Mat& b, int sign) {
    assert(a.rows == b.rows && a.cols == b.cols);
    Mat d(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) {
        if (a.type() == CV_32F) d.at<float>(i, j) = sign > 0 ? a.at<float>(i, j) + b.at<float>(i, j) : a.at<float>(i, j) - b.at<float>(i, j);
        else d.set(i, j, a.get(i, j) + sign * b.get(i, j));
    }
    return d;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return binop(a, b, 1); }
static inline Mat operator-(const Mat& a, const Mat& b) { return binop(a, b, -1); }
static inline Mat operator+(const MatScaleExpr& s, const Mat& b) { return binop(Mat(s), b, 1); }
static inline Mat operator-(const MatScaleExpr& s, const Mat& b) { return binop(Mat(s), b, -1); }

// cv::norm(Mat) NORM_L2: sum of squares in double, sqrt
static inline double norm(const Mat& m) { double s = 0; for (int i = 0; i < m.rows; i++) for (int j = 0; j < m.cols; j++) { const double v = m.get(i, j); s += v * v; } return std::sqrt(s); }

template <class T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) { assert(m.empty() || m.type() == DataType<T>::type); }
    Mat_(const MatInit& e) : Mat(e) { assert(e.type == DataType<T>::type); }
    T& operator()(int r, int c) { return at<T>(r, c); }
    const T& operator()(int r, int c) const { return at<T>(r, c); }
    T& operator()(int i) { return at<T>(i); }
    const T& operator()(int i) const { return at<T>(i); }
};
template <class T> struct MatCommaInitializer_ {
    Mat_<T> m; int idx;
    MatCommaInitializer_(const Mat_<T>& m_) : m(m_), idx(0) {}
    template <class U> MatCommaInitializer_& operator,(U v) { m.template at<T>(idx / m.cols, idx % m.cols) = (T)v; idx++; return *this; }
    operator Mat_<T>() const { return m; }
    operator Mat() const { return m; }
};
template <class T, class U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) { MatCommaInitializer_<T> ci(m); return (ci, v); }

// ---- _InputArray / _OutputArray: the proxy semantics the reference relies on (getMat, create, release, empty)
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    template <class T> _InputArray(const Mat_<T>& m) : m_(const_cast<Mat*>(static_cast<const Mat*>(&m))) {}
    _InputArray(const MatMulExpr& e) : own_(new Mat(e.eval())), m_(own_.get()) {}
    _InputArray(const MatScaleExpr& e) : own_(new Mat(e)), m_(own_.get()) {}
    _InputArray(const MatInit& e) : own_(new Mat(e)), m_(own_.get()) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    std::shared_ptr<Mat> own_;
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(const Mat& m) : _InputArray() { own_.reset(new Mat(m)); m_ = own_.get(); }   // a view: writes land in the shared buffer
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size s, int type) const { m_->create(s, type); }
    void release() const { m_->release(); }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline void Mat::copyTo(const _OutputArray& dst) const { copyTo(dst.getMatRef()); }
static inline _InputArray noArray() { return _InputArray(); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

// ---- primitives (models: oracle/cvprim.h, pinned against cv2 4.13)
static inline float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

static inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    const Mat img = image.getMat();
    assert(img.type() == CV_8UC1 && nonmaxSuppression);
    std::vector<orc::FastKp> k;
    keypoints.clear();
    if (img.rows < 7 || img.cols < 7) return;
    orc::fast_nms(img.data, img.cols, img.rows, (int)img.step, threshold, k);
    keypoints.reserve(k.size());
    for (const orc::FastKp& q : k) keypoints.push_back(KeyPoint((float)q.x, (float)q.y, 7.f, -1.f, (float)q.response));
}

static inline void resize(InputArray src_, OutputArray dst_, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    const Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && fx == 0 && fy == 0);
    dst_.create(dsize, src.type());
    Mat dst = dst_.getMat();
    orc::resize_linear(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar()) {
    const Mat src = src_.getMat();
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && src.type() == CV_8UC1);
    // (the reference always passes a source that either is separate or is the centre ROI of dst with BORDER_ISOLATED: the border is
    // synthesised from the ROI's own pixels in both cases)
    Mat tmp = src.clone();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    for (int y = 0; y < dst.rows; y++) {
        const int sy = orc::reflect101(y - top, tmp.rows);
        for (int x = 0; x < dst.cols; x++) dst.at<uchar>(y, x) = tmp.at<uchar>(sy, orc::reflect101(x - left, tmp.cols));
    }
}

static inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    const Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat tmp(src.rows, src.cols, CV_8UC1);
    orc::gaussian7(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    for (int y = 0; y < src.rows; y++) std::memcpy(dst.ptr(y), tmp.ptr(y), (size_t)src.cols);
}

static inline void remap(InputArray src_, OutputArray dst_, InputArray map1_, InputArray map2_, int interpolation, int borderMode = BORDER_CONSTANT, const Scalar& = Scalar()) {
    const Mat src = src_.getMat(), m1 = map1_.getMat(), m2 = map2_.getMat();
    assert(interpolation == INTER_LINEAR && borderMode == BORDER_CONSTANT && m1.type() == CV_32F && m2.type() == CV_32F);
    dst_.create(m1.rows, m1.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    for (int y = 0; y < dst.rows; y++)
        for (int x = 0; x < dst.cols; x++) dst.at<uchar>(y, x) = orc::remap_pixel(src.data, src.cols, src.rows, (int)src.step, m1.at<float>(y, x), m2.at<float>(y, x));
}

// ---- GUI / IO entry points some out-of-path helper functions of the compiled sources mention (never called on the hot path)
static inline void shim_unavailable(const char* what) { std::fprintf(stderr, "cvshim: %s is not available in the oracle build\n", what); std::abort(); }
static inline void cvtColor(InputArray, OutputArray, int) { shim_unavailable("cvtColor"); }
static inline void drawKeypoints(InputArray, const std::vector<KeyPoint>&, InputOutputArray) { shim_unavailable("drawKeypoints"); }
static inline void imshow(const std::string&, InputArray) { shim_unavailable("imshow"); }
static inline int waitKey(int = 0) { shim_unavailable("waitKey"); return 0; }

class FileNode {
public:
    FileNode operator[](const char*) const { shim_unavailable("FileNode"); return FileNode(); }
    FileNode operator[](const std::string&) const { shim_unavailable("FileNode"); return FileNode(); }
    FileNode operator[](int) const { shim_unavailable("FileNode"); return FileNode(); }
    size_t size() const { return 0; }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator float() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { shim_unavailable("FileStorage"); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv

#pragma GCC pop_options
using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;
#endif
