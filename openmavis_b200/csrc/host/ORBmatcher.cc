// ORBmatcher.cc — array-level ORB_SLAM3::ORBmatcher on top of liborbx.so (Hamming cores of src/ORBmatcher.cc).
#include "ORBmatcher.h"

#include <stdio.h>
#include <stdlib.h>

#include "../../../include/orbx.h"

namespace ORB_SLAM3 {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

static void die(const char *what) {   // no CPU fallback: fail loudly
    fprintf(stderr, "ORBmatcher::%s: %s\n", what, orbx_last_error());
    abort();
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri, int device) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
    if (orbx_matcher_create(&mHandle, device) != ORBX_OK) die("ORBmatcher");
}
ORBmatcher::~ORBmatcher() { orbx_matcher_destroy(mHandle); }

// One pair: 8 popcounts on the host.  This static member is the API's host side (the reference's remaining host callers —
// MapPoint::ComputeDistinctiveDescriptors, LoopClosing — call it per pair): a kernel launch plus two copies per 256-bit popcount
// would be thousands of times slower than the arithmetic.  Every batch path (orbx_descriptor_distance, the matchers) is on the device.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {
    const unsigned int *pa = a.ptr<unsigned int>(), *pb = b.ptr<unsigned int>();   // rows are 4-byte aligned, as in src/ORBmatcher.cc:2578-2579
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

void ORBmatcher::KnnMatch2(const cv::Mat &query, const cv::Mat &train, std::vector<int> &idx, std::vector<int> &dist) {
    idx.assign((size_t)query.rows * 2, -1); dist.assign((size_t)query.rows * 2, -1);
    if (!query.isContinuous() || !train.isContinuous()) { fprintf(stderr, "KnnMatch2: descriptor matrices must be continuous\n"); abort(); }
    if (orbx_knn2(mHandle, query.ptr<unsigned char>(), query.rows, train.ptr<unsigned char>(), train.rows, idx.data(), dist.data()) != ORBX_OK)
        die("KnnMatch2");
}

std::vector<unsigned char> ORBmatcher::RatioTest(const std::vector<int> &dist, double ratio) {
    std::vector<unsigned char> keep(dist.size() / 2);
    orbx_ratio_test(dist.data(), (int)keep.size(), ratio, keep.data());
    return keep;
}

int ORBmatcher::SearchForTriangulation(const std::vector<int> &nodes1, const std::vector<int> &ptr1, const std::vector<int> &idx1,
                                       const std::vector<int> &nodes2, const std::vector<int> &ptr2, const std::vector<int> &idx2,
                                       const cv::Mat &desc1, const cv::Mat &desc2, const std::vector<unsigned char> &skip1,
                                       const std::vector<unsigned char> &skip2, std::vector<std::pair<size_t, size_t>> &vMatchedPairs) {
    std::vector<int> m12((size_t)desc1.rows, -1);
    int nm = 0;
    if (orbx_search_for_triangulation(mHandle, (int)nodes1.size(), nodes1.data(), ptr1.data(), idx1.data(), (int)nodes2.size(),
                                      nodes2.data(), ptr2.data(), idx2.data(), desc1.ptr<unsigned char>(), desc1.rows,
                                      desc2.ptr<unsigned char>(), desc2.rows, skip1.empty() ? nullptr : skip1.data(),
                                      skip2.empty() ? nullptr : skip2.data(), m12.data(), &nm) != ORBX_OK)
        die("SearchForTriangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nm);
    for (size_t i = 0; i < m12.size(); i++)           // src/ORBmatcher.cc:1448-1453: ascending idx1
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
    return nm;
}

int ORBmatcher::MatchCandidates(const cv::Mat &query, const cv::Mat &train, const std::vector<int> &row_ptr, const std::vector<int> &cand,
                                std::vector<int> &bestIdx, std::vector<int> &bestDist, std::vector<int> &secondDist,
                                std::vector<int> &bestPos, std::vector<int> &secondPos) {
    const size_t nq = (size_t)query.rows;
    bestIdx.assign(nq, -1); bestDist.assign(nq, 256); secondDist.assign(nq, 256); bestPos.assign(nq, -1); secondPos.assign(nq, -1);
    if (orbx_match_csr(mHandle, query.ptr<unsigned char>(), query.rows, train.ptr<unsigned char>(), train.rows, row_ptr.data(), cand.data(),
                       bestIdx.data(), bestDist.data(), secondDist.data(), bestPos.data(), secondPos.data()) != ORBX_OK)
        die("MatchCandidates");
    return (int)nq;
}

int ORBmatcher::SearchByBoW(const std::vector<int> &nodes1, const std::vector<int> &ptr1, const std::vector<int> &idx1,
                            const std::vector<int> &nodes2, const std::vector<int> &ptr2, const std::vector<int> &idx2, const cv::Mat &desc1,
                            const cv::Mat &desc2, const std::vector<unsigned char> &skip1, const std::vector<unsigned char> &skip2,
                            const std::vector<float> &angle1, const std::vector<float> &angle2, std::vector<int> &match12) {
    match12.assign((size_t)desc1.rows, -1);
    int nm = 0;
    if (orbx_search_by_bow_keyframes(mHandle, (int)nodes1.size(), nodes1.data(), ptr1.data(), idx1.data(), (int)nodes2.size(), nodes2.data(),
                                     ptr2.data(), idx2.data(), desc1.ptr<unsigned char>(), desc1.rows, desc2.ptr<unsigned char>(), desc2.rows,
                                     skip1.empty() ? nullptr : skip1.data(), skip2.empty() ? nullptr : skip2.data(), angle1.data(),
                                     angle2.data(), mfNNratio, mbCheckOrientation ? 1 : 0, match12.data(), &nm) != ORBX_OK)
        die("SearchByBoW");
    return nm;
}

int ORBmatcher::SearchForInitialization(const std::vector<cv::KeyPoint> &keys1, const cv::Mat &desc1, const std::vector<cv::KeyPoint> &keys2,
                                        const cv::Mat &desc2, float minX, float maxX, float minY, float maxY,
                                        std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint and orbx_kp must have the same layout");
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f must be two packed floats");
    vnMatches12.assign(keys1.size(), -1);                               // :902
    const float gb[4] = {minX, maxX, minY, maxY};
    int nm = 0;
    if (orbx_search_for_initialization(mHandle, (const orbx_kp *)keys1.data(), desc1.ptr<unsigned char>(), (int)keys1.size(),
                                       (const orbx_kp *)keys2.data(), desc2.ptr<unsigned char>(), (int)keys2.size(), gb,
                                       (float *)vbPrevMatched.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, vnMatches12.data(),
                                       &nm) != ORBX_OK)
        die("SearchForInitialization");
    return nm;
}

}  // namespace ORB_SLAM3
