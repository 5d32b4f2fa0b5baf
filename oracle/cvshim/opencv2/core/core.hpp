// oracle/cvshim — TEST INFRASTRUCTURE.  A minimal stand-in for the OpenCV C++ headers, holding exactly what the
// UNMODIFIED reference sources (/root/reference/src/ORBextractor.cc, and the parts of src/ORBmatcher.cc / src/Frame.cc that
// oracle/Makefile lifts at build time) touch, so that they compile in an image that has no OpenCV headers.  Containers
// (Mat, Point_, KeyPoint, InputArray ...) are restated here with OpenCV's semantics for the members used; the imgproc /
// features2d PRIMITIVES (resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2) delegate to the functions of
// oracle/orb_oracle.cpp that tests/test_oracle_vs_cv2.py pins bit-for-bit against the cv2 4.13 wheel — so
// "reference glue + cv2-pinned primitives" is what oracle/_ref/libref_orb.so computes.
// Nothing under openmavis_b200/ includes this directory.
#pragma once
#include <assert.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

// cvRound: OpenCV uses cvtsd2si / cvtss2si (round-half-to-even in the default rounding mode), fast_math.hpp.
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

extern "C" {   // the cv2-pinned primitives of oracle/orb_oracle.cpp
void orc_resize_linear(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride);
void orc_border_reflect101(uint8_t *buf, int w, int h, int stride, int b);
int orc_fast(const uint8_t *img, int w, int h, int stride, int threshold, int *xs, int *ys, int *scores, int cap);
void orc_blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
float orc_fast_atan2(float y, float x);
int orc_hamming(const uint8_t *a, const uint8_t *b);
}

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_L1 = 2, NORM_HAMMING = 6 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class U> Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
};
// types.hpp: a.x = saturate_cast<T>(a.x * b) — for float that is the float product
template <class T> static inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> static inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

// cv::KeyPoint, types.hpp: field order pt, size, angle, response, octave, class_id (28 bytes)
class KeyPoint {
  public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38F) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// cv::Mat for one-channel 8-bit (and, for the matcher stubs, 32-bit float) matrices: reference-counted buffer, ROI views
// that share it, create() that keeps an existing buffer of the right size (what lets cv::resize write into the ROI of
// `temp` at ORBextractor.cc:1083).
class Mat {
  public:
    int rows, cols;
    uchar *data;
    size_t step;   // bytes per row (cv::MatStep converts to size_t)

    Mat() : rows(0), cols(0), data(0), step(0), type_(CV_8U) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), type_(CV_8U) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), data(0), step(0), type_(CV_8U) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void *ext, size_t st = 0)
        : rows(r), cols(c), data((uchar *)ext), step(st ? st : (size_t)c * esz(type)), type_(type) {}

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;   // Mat::create: no reallocation when shape matches
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        size_t n = (size_t)(r > 0 ? r : 0) * step;
        own_.reset(new uchar[n + 64], std::default_delete<uchar[]>());
        data = own_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.rows > 0) memset(m.data, 0, (size_t)m.rows * m.step); return m; }

    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    void release() { own_.reset(); data = 0; rows = cols = 0; step = 0; }
    int type() const { return type_; }
    size_t elemSize() const { return esz(type_); }
    size_t step1() const { return step / esz(type_); }   // single-channel: elements per row
    bool isContinuous() const { return step == (size_t)cols * esz(type_); }
    Size size() const { return Size(cols, rows); }

    template <class T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <class T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <class T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }

    Mat view(int x, int y, int w, int h) const {
        Mat m; m.rows = h; m.cols = w; m.step = step; m.type_ = type_; m.own_ = own_;
        m.data = data + (size_t)y * step + (size_t)x * esz(type_);
        return m;
    }
    Mat operator()(const Rect &r) const { return view(r.x, r.y, r.width, r.height); }
    Mat rowRange(int a, int b) const { return view(0, a, cols, b - a); }
    Mat colRange(int a, int b) const { return view(a, 0, b - a, rows); }
    Mat row(int r) const { return view(0, r, cols, 1); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_));
        return m;
    }
    inline void copyTo(const class _OutputArray &dst) const;
    Mat getMat() const { return *this; }

  private:
    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    int type_;
    std::shared_ptr<uchar> own_;
};

class _InputArray {
  public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
  private:
    const Mat *m_;
};
class _OutputArray {
  public:
    _OutputArray(Mat &m) : m_(&m) {}
    _OutputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}   // OpenCV: fixed-size output (row views, temporaries)
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz, type); }
    Mat getMat() const { return *m_; }
    Mat &getMatRef() const { return *m_; }
  private:
    Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

inline void Mat::copyTo(const _OutputArray &dst) const {
    dst.create(rows, cols, type_);
    Mat d = dst.getMat();
    for (int r = 0; r < rows; r++) memmove(d.ptr(r), ptr(r), (size_t)cols * esz(type_));
}

static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::norm(a, b, NORM_L1) of two 8-bit single-channel matrices of equal size: the sum of absolute differences as a double (OpenCV
// accumulates 8-bit L1 norms in integers and returns the exact sum; tests/test_oracle_vs_cv2.py pins this against cv2.norm).
// Called by Frame::ComputeStereoMatches (src/Frame.cc:1167).
static inline double norm(const Mat &a, const Mat &b, int normType) {
    long long s = 0;
    if (normType != NORM_L1 || a.type() != CV_8U || b.type() != CV_8U || a.rows != b.rows || a.cols != b.cols) return -1.0;
    for (int y = 0; y < a.rows; y++) {
        const uchar *pa = a.data + (size_t)y * a.step, *pb = b.data + (size_t)y * b.step;
        for (int x = 0; x < a.cols; x++) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
    }
    return (double)s;
}

// cv::FileStorage / cv::FileNode: inert stand-ins.  DBoW2's TemplatedVocabulary has VIRTUAL save / load members on them, which are
// instantiated with the class; nothing in the tests calls them (the vocabulary is read with the reference's loadFromTextFile).
class FileNode {
  public:
    size_t size() const { return 0; }
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
  public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
};
template <class T> static inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }

}  // namespace cv
