// ORBextractor.h — header-compatible replacement of the reference's include/ORBextractor.h:28-84, implemented on the C ABI
// of liborbx.so (include/orbx.h).  Same class name, namespace, public signatures and public data member, so
// src/Frame.cc (ExtractORB :599-640, the constructors :211-214, :1320-1337, :1841-1862), src/Tracking.cc:572-587 compile
// unchanged against it.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <memory>
#include <vector>

#include "orbx_cv_compat.h"

struct orbx_extractor;

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // include/ORBextractor.h:33-35, src/ORBextractor.cc:351-414
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image; mask is ignored (src/ORBextractor.cc:987-1071).
    // Returns the number of keypoints outside the lapping area, -1 for an empty image.
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors,
                   std::vector<int> &vLappingArea);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Host mirrors of the pyramid levels of the last call.  The only reader in the reference is Frame::ComputeStereoMatches
    // (src/Frame.cc:1063,1147,1163), which has a device entry point of its own (orbx_compute_stereo_matches reads the device-resident
    // pyramids): the mirrors are 3.7x the image bytes per call over PCIe, so they are NOT downloaded unless switched on.
    std::vector<cv::Mat> mvImagePyramid;
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }
    // extension: CUDA device of this extractor (default 0) — call before the first operator()
    void SetDevice(int device) { mDevice = device; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

private:
    std::shared_ptr<void> mRig;        // rig context shared by the extractors of one geometry (one batch-capable device handle)
    int mWidth = 0, mHeight = 0, mDevice = 0;
    bool mbDownloadPyramid = false;
    std::vector<unsigned char> mKpBuf, mDescBuf;
};

}  // namespace ORB_SLAM3
#endif
