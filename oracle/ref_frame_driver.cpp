// oracle/ref_frame_driver.cpp — TEST INFRASTRUCTURE.  C entry points around two reference functions of SURVEY.md §8(f) whose BODIES are
// compiled unmodified: oracle/Makefile lifts them by line range from the files where they lie under /root/reference into
// oracle/_ref/gen/ at build time (never into the repo) and this file includes them:
//   Frame::ComputeStereoMatches                (src/Frame.cc:1057-1219)        row N2  -> _ref/gen/frame_stereo.inc
//   MapPoint::ComputeDistinctiveDescriptors    (src/MapPoint.cc:405-483)       row N4  -> _ref/gen/mappoint_distinctive.inc
//   ORBmatcher::DescriptorDistance             (src/ORBmatcher.cc:2575-2591)   (called by both) -> _ref/gen/orbmatcher_distance.inc
// What stands in for the rest of ORB-SLAM3 are the classes below: exactly the data members the two bodies touch, with the reference's
// names (include/Frame.h, include/KeyFrame.h, include/MapPoint.h, include/ORBextractor.h:86 mvImagePyramid).  cv::Mat / cv::KeyPoint /
// cv::norm come from oracle/cvshim.  tests/test_ref_frame.py requires reference == oracle (orc_compute_stereo_matches,
// orc_distinctive_descriptors), which is what the device paths orbx_compute_stereo_matches / orbx_distinctive_descriptors are tested against.
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

using namespace std;   // the reference's sources name vector / pair / map / tuple / unique_lock unqualified

namespace ORB_SLAM3 {

class ORBmatcher {   // include/ORBmatcher.h:24-26, 29
  public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
};
const int ORBmatcher::TH_HIGH = 100;   // src/ORBmatcher.cc:16-17
const int ORBmatcher::TH_LOW = 50;

class ORBextractor {   // include/ORBextractor.h:86
  public:
    std::vector<cv::Mat> mvImagePyramid;
};

class Frame {
  public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mb = 0, mbf = 0;
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    void ComputeStereoMatches();
};

class KeyFrame {
  public:
    cv::Mat mDescriptors;
    bool bad = false;
    bool isBad() { return bad; }
};

class MapPoint {
  public:
    std::map<KeyFrame *, std::tuple<int, int, int, int>> mObservations;   // include/MapPoint.h: (left, right, side-left, side-right) indices
    std::mutex mMutexFeatures;
    bool mbBad = false;
    cv::Mat mDescriptor;
    void ComputeDistinctiveDescriptors();
};

#include "_ref/gen/orbmatcher_distance.inc"
#include "_ref/gen/frame_stereo.inc"
#include "_ref/gen/mappoint_distinctive.inc"

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

struct ref_kp { float x, y, size, angle, response; int octave, class_id; };   // = orbx_kp / the oracle's KeyPt

static void fill_keys(std::vector<cv::KeyPoint> &out, const ref_kp *k, int n) {
    out.resize(n);
    for (int i = 0; i < n; i++) {
        out[i].pt.x = k[i].x; out[i].pt.y = k[i].y; out[i].size = k[i].size; out[i].angle = k[i].angle; out[i].response = k[i].response;
        out[i].octave = k[i].octave; out[i].class_id = k[i].class_id;
    }
}

extern "C" {

// Frame::ComputeStereoMatches on a rectified pair.  Keypoints / descriptors of both images, the scale tables, and both image pyramids:
// level l of the left image = lw[l] x lh[l] pixels at pyrL[l] with row stride lstride[l] (the level interior: in the reference
// mvImagePyramid[l] is a view into the bordered buffer, src/ORBextractor.cc:1079-1082, and so it is here).  Outputs mvuRight[nl], mvDepth[nl].
// Returns the number of keypoints that kept a depth.
int reff_compute_stereo_matches(const ref_kp *kl, const uint8_t *dl, int nl, const ref_kp *kr, const uint8_t *dr, int nr, const float *sf,
                                const float *inv_sf, int nlevels, const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *lw,
                                const int *lh, const int *lstride, float mb, float mbf, float *mvuRight, float *mvDepth) {
    Frame F;
    ORBextractor EL, ER;
    F.N = nl;
    fill_keys(F.mvKeys, kl, nl); fill_keys(F.mvKeysRight, kr, nr);
    F.mDescriptors = cv::Mat(nl, 32, CV_8U, (void *)dl); F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U, (void *)dr);
    F.mvScaleFactors.assign(sf, sf + nlevels); F.mvInvScaleFactors.assign(inv_sf, inv_sf + nlevels);
    F.mb = mb; F.mbf = mbf;
    for (int l = 0; l < nlevels; l++) {
        EL.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void *)pyrL[l], (size_t)lstride[l]));
        ER.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void *)pyrR[l], (size_t)lstride[l]));
    }
    F.mpORBextractorLeft = &EL; F.mpORBextractorRight = &ER;
    F.ComputeStereoMatches();
    int kept = 0;
    for (int i = 0; i < nl; i++) { mvuRight[i] = F.mvuRight[i]; mvDepth[i] = F.mvDepth[i]; kept += F.mvDepth[i] > 0; }
    return kept;
}

// MapPoint::ComputeDistinctiveDescriptors for `ngroups` map points.  Map point g is observed by nkf[g] keyframes (at most 16); keyframe j
// of point g contributes the descriptor rows idx[(ptr4[g] + j) * 4 + c] (c = left, right, side-left, side-right; -1 = not observed)
// of its own descriptor matrix, which is rows [kf_row0[ptr4[g] + j], ...) of `desc`; bad[...] marks keyframes with isBad().
// std::map iterates keyframes in POINTER order: the driver allocates the keyframes of a point in one array, so that order = j.
// best_row[g] = the row of `desc` whose bytes ended up in mDescriptor, or -1 when the function returned early.
void reff_distinctive_descriptors(const uint8_t *desc, const int *ptr4, const int *kf_row0, const int *kf_rows, const int *idx, const uint8_t *bad,
                                  int ngroups, int mp_bad_every, int *best_row) {
    for (int g = 0; g < ngroups; g++) {
        const int n = ptr4[g + 1] - ptr4[g];
        std::vector<KeyFrame> kfs(n);
        MapPoint P;
        for (int j = 0; j < n; j++) {
            const int q = ptr4[g] + j;
            kfs[j].mDescriptors = cv::Mat(kf_rows[q], 32, CV_8U, (void *)(desc + (size_t)kf_row0[q] * 32));
            kfs[j].bad = bad[q] != 0;
            P.mObservations[&kfs[j]] = std::make_tuple(idx[4 * q], idx[4 * q + 1], idx[4 * q + 2], idx[4 * q + 3]);
        }
        P.mbBad = mp_bad_every > 0 && (g % mp_bad_every) == mp_bad_every - 1;
        P.ComputeDistinctiveDescriptors();
        best_row[g] = -1;
        if (P.mDescriptor.empty()) continue;
        // identify the chosen row by address arithmetic is impossible after clone(): compare bytes against the candidates in push order
        for (int j = 0; j < n && best_row[g] < 0; j++) {
            const int q = ptr4[g] + j;
            if (bad[q]) continue;
            for (int c = 0; c < 4 && best_row[g] < 0; c++)
                if (idx[4 * q + c] != -1 && memcmp(P.mDescriptor.data, desc + (size_t)(kf_row0[q] + idx[4 * q + c]) * 32, 32) == 0)
                    best_row[g] = kf_row0[q] + idx[4 * q + c];
        }
    }
}

// the stand-in cv::norm(a, b, NORM_L1) itself (oracle/cvshim), pinned against cv2.norm by tests/test_ref_frame.py
double reff_norm_l1(const uint8_t *a, int astride, const uint8_t *b, int bstride, int rows, int cols) {
    return cv::norm(cv::Mat(rows, cols, CV_8U, (void *)a, (size_t)astride), cv::Mat(rows, cols, CV_8U, (void *)b, (size_t)bstride), cv::NORM_L1);
}

int reff_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    cv::Mat ma(1, 32, CV_8U, (void *)a), mb(1, 32, CV_8U, (void *)b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

}  // extern "C"
