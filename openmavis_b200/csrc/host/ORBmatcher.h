// ORBmatcher.h — the array-level matcher surface of the reference's include/ORBmatcher.h:17-123 on the C ABI of liborbx.so.
//
// The reference's ORBmatcher methods take Frame / KeyFrame / MapPoint graph objects; those classes (and their mutexes,
// covisibility graph, g2o types) are out of scope.  What crosses into liborbx is the Hamming core of each
// method.  This class keeps the reference's constants, constructor and static DescriptorDistance, and exposes every core
// with plain-array arguments; INTEGRATION.md shows, method by method, the few lines a maintainer adds inside
// src/ORBmatcher.cc to marshal Frame/KeyFrame members into these calls.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <utility>
#include <vector>

#include "orbx_cv_compat.h"

struct orbx_matcher;

namespace ORB_SLAM3 {

class ORBmatcher {
public:
    static const int TH_LOW;         // 50   src/ORBmatcher.cc:15
    static const int TH_HIGH;        // 100  :16
    static const int HISTO_LENGTH;   // 30   :17

    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0);   // include/ORBmatcher.h:21
    ~ORBmatcher();
    ORBmatcher(const ORBmatcher &) = delete;              // owns a device handle
    ORBmatcher &operator=(const ORBmatcher &) = delete;

    // Computes the Hamming distance between two ORB descriptors (src/ORBmatcher.cc:2577-2591).  Each row: 32 bytes.
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    // cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2) + the caller's ratio test (src/Frame.cc:1427-1434,
    // 1483-1491): idx/dist [nq][2] ordered (distance, trainIdx); keep[i] = d0 < d1 * ratio.
    void KnnMatch2(const cv::Mat &query, const cv::Mat &train, std::vector<int> &idx, std::vector<int> &dist);
    static std::vector<unsigned char> RatioTest(const std::vector<int> &dist, double ratio);

    // Hamming core of SearchForTriangulation (src/ORBmatcher.cc:1131-1456, bCoarse): FeatureVectors as CSR.
    int SearchForTriangulation(const std::vector<int> &nodes1, const std::vector<int> &ptr1, const std::vector<int> &idx1,
                               const std::vector<int> &nodes2, const std::vector<int> &ptr2, const std::vector<int> &idx2,
                               const cv::Mat &desc1, const cv::Mat &desc2, const std::vector<unsigned char> &skip1,
                               const std::vector<unsigned char> &skip2, std::vector<std::pair<size_t, size_t>> &vMatchedPairs);

    // Candidate-list core shared by SearchByProjection(Frame,Frame / Frame,KF / KF,Sim3), Fuse, SearchBySim3,
    // SearchForInitialization (src/ORBmatcher.cc:72-114 and clones): best / second per query over cand[row_ptr[q]..).
    int MatchCandidates(const cv::Mat &query, const cv::Mat &train, const std::vector<int> &row_ptr, const std::vector<int> &cand,
                        std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist, std::vector<int> &bestPos,
                        std::vector<int> &secondPos);

    // ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:1006-1129), whole function:
    // FeatureVectors as CSR, skip = feature without usable MapPoint, angles = mvKeysUn[i].angle.  match12[idx1] = idx2 or -1.
    int SearchByBoW(const std::vector<int> &nodes1, const std::vector<int> &ptr1, const std::vector<int> &idx1,
                    const std::vector<int> &nodes2, const std::vector<int> &ptr2, const std::vector<int> &idx2, const cv::Mat &desc1,
                    const cv::Mat &desc2, const std::vector<unsigned char> &skip1, const std::vector<unsigned char> &skip2,
                    const std::vector<float> &angle1, const std::vector<float> &angle2, std::vector<int> &match12);

    // ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:895-1004), whole function, on the two frames' mvKeysUn / mDescriptors.
    int SearchForInitialization(const std::vector<cv::KeyPoint> &keys1, const cv::Mat &desc1, const std::vector<cv::KeyPoint> &keys2,
                                const cv::Mat &desc2, float minX, float maxX, float minY, float maxY,
                                std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10);

    // The C-ABI handle: every other matcher (orbx_search_by_projection, orbx_search_by_projection_last, orbx_search_by_bow_frame,
    // orbx_window_best, orbx_search_by_sim3, orbx_stereo_fisheye_matches, ... — include/orbx.h) takes it as its first argument.
    orbx_matcher *Handle() const { return mHandle; }

    float mfNNratio;
    bool mbCheckOrientation;

private:
    orbx_matcher *mHandle = nullptr;
};

}  // namespace ORB_SLAM3
#endif
