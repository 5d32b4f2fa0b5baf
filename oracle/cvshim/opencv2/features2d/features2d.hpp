// oracle/cvshim — test infrastructure, see opencv2/core/core.hpp.  cv::FAST (TYPE_9_16) on the cv2-pinned oracle primitive,
// KeyPointsFilter::retainBest (only reachable from the reference's unused ComputeKeyPointsOld), BFMatcher 2-NN.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "opencv2/core/core.hpp"

namespace cv {

// cv::FAST(image, keypoints, threshold, nonmaxSuppression) — fast.cpp: KeyPoint(x, y, 7.f, -1, score)
static inline void FAST(InputArray img_, std::vector<KeyPoint> &kps, int threshold, bool nonmax = true) {
    Mat img = img_.getMat();
    kps.clear();
    if (!nonmax) { fprintf(stderr, "cvshim: FAST without NMS is not pinned\n"); abort(); }
    if (img.empty()) return;
    int cap = img.rows * img.cols;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    int n = orc_fast(img.data, img.cols, img.rows, (int)img.step, threshold, xs.data(), ys.data(), sc.data(), cap);
    kps.reserve(n);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

struct KeyPointsFilter {
    // keypoint.cpp: keep the n best by response, plus everything that ties with the n-th
    static void retainBest(std::vector<KeyPoint> &kps, int n) {
        if (n >= 0 && kps.size() > (size_t)n) {
            if (n == 0) { kps.clear(); return; }
            std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(),
                             [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
            float amb = kps[n - 1].response;
            auto it = std::partition(kps.begin() + n, kps.end(), [amb](const KeyPoint &k) { return k.response >= amb; });
            kps.resize(it - kps.begin());
        }
    }
};

}  // namespace cv
