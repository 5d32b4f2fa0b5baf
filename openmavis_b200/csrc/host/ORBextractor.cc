// ORBextractor.cc — ORB_SLAM3::ORBextractor on top of liborbx.so (replaces src/ORBextractor.cc of the reference).
#include "ORBextractor.h"

#include <stdio.h>
#include <stdlib.h>

#include "../../../include/orbx.h"

namespace ORB_SLAM3 {

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    // the per-level tables do not depend on the image size: a 1-image probe handle is not needed, liborbx computes them
    // from (scaleFactor, nlevels) exactly like src/ORBextractor.cc:362-376 when the first handle is created; until then the
    // getters are served from the same arithmetic below.
    mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_extractor_destroy(mHandle); }

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                             cv::OutputArray _descriptors, std::vector<int> &vLappingArea) {
    if (_image.empty()) return -1;                       // src/ORBextractor.cc:994-995
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) { fprintf(stderr, "ORBextractor: image.type() must be CV_8UC1\n"); abort(); }   // :998 assert
    if (!mHandle || image.cols != mWidth || image.rows != mHeight) {
        orbx_extractor_destroy(mHandle);
        mHandle = nullptr;
        if (orbx_extractor_create(&mHandle, mDevice, image.cols, image.rows, 1, nfeatures, (float)scaleFactor, nlevels, 0) != ORBX_OK) {
            fprintf(stderr, "ORBextractor: %s\n", orbx_last_error());   // no CPU fallback: fail loudly
            abort();
        }
        mWidth = image.cols; mHeight = image.rows;
        orbx_extractor_tables(mHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                              nullptr, nullptr);
    }
    const int cap = orbx_extractor_capacity(mHandle);
    mKpBuf.resize((size_t)cap * sizeof(orbx_kp));
    mDescBuf.resize((size_t)cap * 32);
    orbx_image_params prm;
    prm.nfeatures = nfeatures; prm.ini_th_fast = iniThFAST; prm.min_th_fast = minThFAST;
    prm.lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0;
    prm.lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
    int n = 0;
    const int mono = orbx_extract(mHandle, image.data, (int)image.step, &prm, (orbx_kp *)mKpBuf.data(), mDescBuf.data(), cap, &n);
    if (mono < 0) { fprintf(stderr, "ORBextractor: %s\n", orbx_last_error()); abort(); }
    // callee allocates into the caller's containers (src/ORBextractor.cc:1013-1022)
    _keypoints = std::vector<cv::KeyPoint>(n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint layout");
    if (n) memcpy((void *)_keypoints.data(), mKpBuf.data(), (size_t)n * sizeof(orbx_kp));
    _descriptors.create(n, 32, CV_8U);
    cv::Mat descriptors = _descriptors.getMat();
    for (int i = 0; i < n; i++) memcpy(descriptors.ptr<unsigned char>(i), mDescBuf.data() + (size_t)i * 32, 32);
    if (mbDownloadPyramid) {
        std::vector<int> lw(nlevels), lh(nlevels);
        orbx_extractor_tables(mHandle, nullptr, nullptr, nullptr, nullptr, lw.data(), lh.data());
        for (int l = 0; l < nlevels; l++) {
            // like the reference, a level is the interior ROI of a (w+38) x (h+38) buffer with the REFLECT_101 frame
            cv::Mat temp(lh[l] + 38, lw[l] + 38, CV_8UC1);
            orbx_extractor_pyramid_level(mHandle, 0, l, 1, temp.data, (int)temp.step);
#ifdef ORBX_HAVE_OPENCV
            mvImagePyramid[l] = temp(cv::Rect(19, 19, lw[l], lh[l]));
#else
            mvImagePyramid[l] = temp.roi(19, 19, lw[l], lh[l]);
#endif
        }
    }
    return mono;
}

}  // namespace ORB_SLAM3
