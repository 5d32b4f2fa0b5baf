// oracle/cvshim — test infrastructure, see opencv2/core/core.hpp.  cv::resize / copyMakeBorder / GaussianBlur for the one
// configuration each that the reference uses; anything else aborts loudly instead of computing something unpinned.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "opencv2/core/core.hpp"

namespace cv {

static inline void cvshim_require(bool ok, const char *what) {
    if (!ok) { fprintf(stderr, "cvshim: unsupported call: %s\n", what); abort(); }
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (ORBextractor.cc:1083); dst keeps its buffer when it already
// has dsize (the pyramid level is an ROI of `temp`)
static inline void resize(InputArray src_, OutputArray dst_, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR) {
    Mat src = src_.getMat();
    cvshim_require(interp == INTER_LINEAR && fx == 0 && fy == 0 && src.type() == CV_8UC1, "resize: INTER_LINEAR, dsize, CV_8UC1 only");
    dst_.create(dsize.height, dsize.width, CV_8UC1);
    Mat dst = dst_.getMat();
    orc_resize_linear(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

// cv::copyMakeBorder(src, dst, t, b, l, r, BORDER_REFLECT_101 [+ BORDER_ISOLATED]) (ORBextractor.cc:1085-1101).  Without
// BORDER_ISOLATED OpenCV reads real pixels around a sub-matrix src; the reference passes whole images there (level 0), for
// which both variants are the same operation — sub-matrix sources are treated as isolated here.
static inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
    Mat src = src_.getMat();
    cvshim_require((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && top == bottom && left == right && top == left && src.type() == CV_8UC1,
                   "copyMakeBorder: equal REFLECT_101 borders, CV_8UC1 only");
    dst_.create(src.rows + 2 * top, src.cols + 2 * left, CV_8UC1);
    Mat dst = dst_.getMat();
    for (int r = 0; r < src.rows; r++) {
        uchar *d = dst.ptr(r + top) + left;
        if (d != src.ptr(r)) memmove(d, src.ptr(r), (size_t)src.cols);
    }
    orc_border_reflect101(dst.data, src.cols, src.rows, (int)dst.step, top);
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (ORBextractor.cc:1036), in place allowed
static inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT) {
    Mat src = src_.getMat();
    cvshim_require(ksize.width == 7 && ksize.height == 7 && sx == 2 && (sy == 2 || sy == 0) && borderType == BORDER_REFLECT_101 && src.type() == CV_8UC1,
                   "GaussianBlur: 7x7, sigma 2, REFLECT_101, CV_8UC1 only");
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    orc_blur7(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);   // blur7 buffers the horizontal pass: in-place safe
}

}  // namespace cv
