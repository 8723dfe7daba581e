// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for the OpenCV 3.4 headers the reference's
// src/orb_extractor.cpp, src/matcher.cpp and src/camera.cpp include (OpenCV is not in this image).
// Only the types and members those three translation units touch exist here; the five image primitives
// (FAST, resize, copyMakeBorder, GaussianBlur, fastAtan2) are implemented in ../cv_standin.cpp by the
// closed-form models of oracle/orb_oracle.c, which tests/test_oracle_vs_cv2.py pins bit-for-bit to cv2.
// The reference sources themselves are compiled UNMODIFIED from /root/reference (see ../Makefile).
#ifndef REF_STANDIN_OPENCV_CORE_HPP_
#define REF_STANDIN_OPENCV_CORE_HPP_
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

// cvRound / cvFloor / cvCeil: OpenCV 3.4 core/fast_math.hpp — cvtsd2si / cvtss2si, i.e. round half to even.
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_ &operator*=(T s) { x = x * s; y = y * s; return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

class KeyPoint {  // 28-byte POD, as in OpenCV
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    MatStep(size_t s) : v(s) {}
    operator size_t() const { return v; }
};

struct MatExpr { int rows, cols, type; };  // only Mat::zeros

class Mat {
public:
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(Size sz, int t) { create(sz.height, sz.width, t); }
    Mat(int r, int c, int t, void *d, size_t s = 0) : data((uchar *)d), rows(r), cols(c), step(s ? s : (size_t)c * esz(t)), type_(t) {}
    Mat(const MatExpr &e) { *this = e; }
    // OpenCV: assigning zeros() to a matrix of the same size and type fills it IN PLACE (Mat::create is a no-op then) —
    // computeDescriptors (src/orb_extractor.cpp:1037) relies on it: `descriptors` is a rowRange view of the output.
    Mat &operator=(const MatExpr &e) {
        create(e.rows, e.cols, e.type);
        for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, (size_t)cols * esz(type_));
        return *this;
    }
    static MatExpr zeros(int r, int c, int t) { MatExpr e = {r, c, t}; return e; }
    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == type_) return;
        own_ = std::shared_ptr<uchar>((uchar *)std::malloc((size_t)r * c * esz(t) + 64), std::free);
        data = own_.get(); rows = r; cols = c; type_ = t; step = (size_t)c * esz(t);
    }
    void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step / esz(type_); }
    uchar *ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar *ptr(int i = 0) const { return data + (size_t)i * step; }
    template <class T> T *ptr(int i = 0) { return (T *)(data + (size_t)i * step); }
    template <class T> const T *ptr(int i = 0) const { return (const T *)(data + (size_t)i * step); }
    template <class T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <class T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
    Mat operator()(const Rect &roi) const {
        Mat m(*this);
        m.data = data + (size_t)roi.y * step + (size_t)roi.x * esz(type_);
        m.rows = roi.height; m.cols = roi.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_));
        return m;
    }
    void convertTo(Mat &, int) const { throw std::logic_error("cv stand-in: convertTo is not on the path"); }

    uchar *data = nullptr;
    int rows = 0, cols = 0;
    MatStep step;

private:
    static size_t esz(int t) { return t == CV_32F ? 4 : 1; }
    int type_ = CV_8UC1;
    std::shared_ptr<uchar> own_;
};

// InputArray / OutputArray: thin views of one Mat (the only kind of array the path passes).
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat &m) { m_ = &m; }
    void release() const { if (m_) m_->release(); }
    void create(int r, int c, int t) const { if (m_) m_->create(r, c, t); }
    void create(Size sz, int t) const { if (m_) m_->create(sz.height, sz.width, t); }
    Mat &getMatRef() const { return *m_; }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
const _OutputArray &noArray();

struct KeyPointsFilter {  // only ComputeKeyPointsOld (dead code, src/orb_extractor.cpp:855-1032) calls it
    static void retainBest(std::vector<KeyPoint> &keypoints, int npoints);
};

// -- the five primitives (cv_standin.cpp) -----------------------------------------------------------------------------
void FAST(InputArray image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
float fastAtan2(float y, float x);
int borderInterpolate(int p, int len, int borderType);

// every FAST() call of the current thread, in call order, when recording is on (the _ref shim reads the cell results)
struct FastCall { const uchar *data; size_t step; int cols, rows, threshold; std::vector<KeyPoint> keypoints; };
std::vector<FastCall> &fastCallLog();
void fastCallLogEnable(bool on);

}  // namespace cv
#endif
