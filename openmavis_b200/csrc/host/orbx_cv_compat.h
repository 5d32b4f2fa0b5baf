// orbx_cv_compat.h — the handful of OpenCV types the ORBextractor / ORBmatcher interface mentions, for builds WITHOUT
// OpenCV (this image has no OpenCV headers).  With OpenCV available compile with -DORBX_HAVE_OPENCV and the real
// <opencv2/core.hpp> is used instead; the shim sources only touch members that both provide.
#pragma once
#ifdef ORBX_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include <stdint.h>
#include <string.h>

#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point { int x = 0, y = 0; };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
// field order and size of cv::KeyPoint (28 bytes)
struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t *data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void *ext, size_t st) : rows(r), cols(c), data((uint8_t *)ext), step(st) {}
    void create(int r, int c, int /*type*/) {
        rows = r; cols = c; step = (size_t)c;
        own_.reset(new uint8_t[(size_t)(r > 0 ? r : 0) * (c > 0 ? c : 0) + 1], std::default_delete<uint8_t[]>());
        data = own_.get();
    }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    int type() const { return CV_8UC1; }
    bool isContinuous() const { return step == (size_t)cols; }
    template <class T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <class T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.data = data + (size_t)r * step; m.step = step; m.own_ = own_; return m; }
    Mat roi(int x, int y, int w, int h) const { Mat m; m.rows = h; m.cols = w; m.data = data + (size_t)y * step + x; m.step = step; m.own_ = own_; return m; }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<uint8_t> own_;
};
typedef const Mat &InputArray;
typedef Mat &OutputArray;
}  // namespace cv
#endif
