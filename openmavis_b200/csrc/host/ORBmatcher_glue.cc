// ORBmatcher_glue.cc — ORB_SLAM3::ORBmatcher methods with the REFERENCE's signatures (include/ORBmatcher.h:17-123, unmodified) on the
// C ABI of liborbx.so: the marshalling a maintainer puts in place of the corresponding bodies of src/ORBmatcher.cc.
//
// Compiled against the integrator's own headers (ORBmatcher.h, Frame.h, KeyFrame.h, MapPoint.h): it only touches the members the
// reference bodies touch.  Every search runs on the device; what stays here is what the reference does around it on the host —
// reading Frame / KeyFrame / MapPoint members into flat arrays and writing the matches back into F.mvpMapPoints / the output vectors.
// Every public method of the class is here.
//
//   ORBmatcher::DescriptorDistance                                   src/ORBmatcher.cc:2577-2591   host popcount (8 words)
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ...)  src/ORBmatcher.cc:23-340      orbx_search_by_projection
//   ORBmatcher::SearchForInitialization(F1, F2, ...)                 src/ORBmatcher.cc:895-1004    orbx_search_for_initialization
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)                  src/ORBmatcher.cc:349-666     orbx_search_by_bow_frame
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)               src/ORBmatcher.cc:1006-1129   orbx_search_by_bow_keyframes
//   ORBmatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)          :1985-2413   orbx_search_by_projection_last
//   ORBmatcher::SearchByProjection(Frame &Cur, KeyFrame*, sAlreadyFound, th, ORBdist) :2415-2535   orbx_search_by_projection_last
//   ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming, cameraID)              :668-777   "
//   ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratio, cameraID) :779-893   "
//   ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th, cameraID)           src/ORBmatcher.cc:1458-1647   orbx_window_best + the graph updates, in point order
//   ORBmatcher::Fuse(KeyFrame*, Sim3f&, vpPoints, th, vpReplacePoint) src/ORBmatcher.cc:1648-1769  orbx_window_best per camera + the updates
//   ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th)       src/ORBmatcher.cc:1771-1983   orbx_search_by_sim3
//   ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse)  :1131-1456   orbx_search_for_triangulation_full
// The two projection matchers compute their projections here, on the host, with the integrator's own Eigen / Sophus types and the
// reference's expressions (a few thousand 3-D points per call); window search, Hamming distances, ordered claims and the rotation
// filter run on the device.
//
// tests/host_harness builds this file against stub Frame / KeyFrame / MapPoint headers that carry the reference's member names and
// tests/test_matcher_glue.py compares it, call by call, with the reference's own object code (the unmodified ORBmatcher.cc compiled against the same stubs).
#include <ORBmatcher.h>   // the REFERENCE's include/ORBmatcher.h from the include path (not the array-level header next to this file)
#ifndef ORBX_GLUE_BUILD
#include <CameraModels/KannalaBrandt8.h>   // GetPrecision() of the fisheye model (SearchForTriangulation)
#endif

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <set>
#include <tuple>
#include <utility>
#include <vector>

#include "orbx.h"

namespace ORB_SLAM3 {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

[[noreturn]] void die(const char *what) {   // no CPU fallback: fail loudly
    fprintf(stderr, "ORBmatcher::%s: %s\n", what, orbx_last_error());
    abort();
}

// ORBmatcher objects are stack objects created at every call site (src/Tracking.cc:2828 ...): the device handle (stream, events,
// staging buffers) is per host thread instead, created on first use.  ORBX_DEVICE selects the GPU of the process (default 0).
orbx_matcher *handle() {
    thread_local orbx_matcher *h = nullptr;
    if (!h) {
        const char *d = getenv("ORBX_DEVICE");
        if (orbx_matcher_create(&h, d ? atoi(d) : 0) != ORBX_OK) die("orbx_matcher_create");
    }
    return h;
}

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint and orbx_kp must have the same layout");

// keypoints of a Frame in global feature order: [mvKeysUn] (Nleft == -1) or [mvKeys | mvKeysRight | mvKeysSideLeft | mvKeysSideRight]
int frame_layout(const Frame &F, std::vector<orbx_kp> &kps, int cam_off[5]) {
    auto add = [&](const std::vector<cv::KeyPoint> &v, int n) { kps.insert(kps.end(), (const orbx_kp *)v.data(), (const orbx_kp *)v.data() + n); };
    kps.clear();
    cam_off[0] = 0;
    if (F.Nleft == -1) {
        add(F.mvKeysUn, F.N);
        for (int c = 1; c <= 4; c++) cam_off[c] = F.N;
        return 1;
    }
    add(F.mvKeys, F.Nleft); cam_off[1] = F.Nleft;
    add(F.mvKeysRight, F.Nright); cam_off[2] = F.Nleft + F.Nright;
    if (F.Nsideleft == -1 && F.Nsideright == -1) { cam_off[3] = cam_off[4] = cam_off[2]; return 2; }
    add(F.mvKeysSideLeft, F.Nsideleft); cam_off[3] = cam_off[2] + F.Nsideleft;
    add(F.mvKeysSideRight, F.Nsideright); cam_off[4] = cam_off[3] + F.Nsideright;
    return 4;
}

// descriptor rows as one contiguous N x 32 block (cv::Mat rows may be strided)
const unsigned char *desc_rows(const cv::Mat &m, std::vector<unsigned char> &tmp) {
    if (m.rows == 0) { tmp.assign(32, 0); return tmp.data(); }
    if (m.isContinuous()) return m.ptr<unsigned char>(0);
    tmp.resize((size_t)m.rows * 32);
    for (int r = 0; r < m.rows; r++) memcpy(&tmp[(size_t)r * 32], m.ptr<unsigned char>(r), 32);
    return tmp.data();
}

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR
void feature_vector_csr(const DBoW2::FeatureVector &fv, std::vector<int> &nodes, std::vector<int> &ptr, std::vector<int> &idx) {
    nodes.clear(); idx.clear(); ptr.assign(1, 0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        nodes.push_back((int)it->first);
        for (size_t k = 0; k < it->second.size(); k++) idx.push_back((int)it->second[k]);
        ptr.push_back((int)idx.size());
    }
}

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// One pair: eight popcounts on the host (the remaining host callers — MapPoint::ComputeDistinctiveDescriptors, LoopClosing — call this
// per pair; a kernel launch per 256-bit popcount would be absurd).  Batches go through orbx_descriptor_distance / the matchers.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {
    const unsigned int *pa = a.ptr<unsigned int>(), *pb = b.ptr<unsigned int>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

// src/ORBmatcher.cc:23-340
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th, const bool bFarPoints,
                                   const float thFarPoints) {
    std::vector<orbx_kp> kps;
    int cam_off[5];
    const int ncam = frame_layout(F, kps, cam_off);
    const int N = F.N, nmp = (int)vpMapPoints.size();
    if (N == 0 || nmp == 0) return 0;
    std::vector<unsigned char> dtmp;
    const unsigned char *desc = desc_rows(F.mDescriptors, dtmp);
    float bounds[16];
    for (int c = 0; c < 4; c++) { bounds[4 * c] = F.mnMinX; bounds[4 * c + 1] = F.mnMaxX; bounds[4 * c + 2] = F.mnMinY; bounds[4 * c + 3] = F.mnMaxY; }
    std::vector<unsigned char> occupied(N, 0);
    for (int i = 0; i < N; i++) occupied[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;     // :77-79
    std::vector<unsigned char> in_view((size_t)nmp * 4, 0), bad(nmp, 0), has_obs(nmp, 0), mp_desc((size_t)nmp * 32, 0);
    std::vector<float> px((size_t)nmp * 4, 0.f), py((size_t)nmp * 4, 0.f), vc((size_t)nmp * 4, 0.f), pxr(nmp, 0.f), depth(nmp, 0.f);
    std::vector<int> level((size_t)nmp * 4, 0);
    for (int m = 0; m < nmp; m++) {
        MapPoint *p = vpMapPoints[m];
        in_view[4 * m] = p->mbTrackInView; in_view[4 * m + 1] = p->mbTrackInViewR; in_view[4 * m + 2] = p->mbTrackInViewSL; in_view[4 * m + 3] = p->mbTrackInViewSR;
        px[4 * m] = p->mTrackProjX; px[4 * m + 1] = p->mTrackProjXR; px[4 * m + 2] = p->mTrackProjXSL; px[4 * m + 3] = p->mTrackProjXSR;
        py[4 * m] = p->mTrackProjY; py[4 * m + 1] = p->mTrackProjYR; py[4 * m + 2] = p->mTrackProjYSL; py[4 * m + 3] = p->mTrackProjYSR;
        level[4 * m] = p->mnTrackScaleLevel; level[4 * m + 1] = p->mnTrackScaleLevelR; level[4 * m + 2] = p->mnTrackScaleLevelSL;
        level[4 * m + 3] = p->mnTrackScaleLevelSR;
        vc[4 * m] = p->mTrackViewCos; vc[4 * m + 1] = p->mTrackViewCosR; vc[4 * m + 2] = p->mTrackViewCosSL; vc[4 * m + 3] = p->mTrackViewCosSR;
        pxr[m] = p->mTrackProjXR; depth[m] = p->mTrackDepth;
        bad[m] = p->isBad(); has_obs[m] = p->Observations() > 0;
        if (in_view[4 * m] | in_view[4 * m + 1] | in_view[4 * m + 2] | in_view[4 * m + 3]) {
            const cv::Mat d = p->GetDescriptor();
            memcpy(&mp_desc[(size_t)m * 32], d.ptr<unsigned char>(0), 32);
        }
    }
    std::vector<int> assigned(N, -1);
    int nmatches = 0;
    const bool stereo = ncam == 1 && (int)F.mvuRight.size() == N;
    if (orbx_search_by_projection(handle(), ncam, cam_off, kps.data(), desc, bounds, F.mvScaleFactors.data(), (int)F.mvScaleFactors.size(),
                                  occupied.data(), ncam >= 2 ? F.mvLeftToRightMatch.data() : nullptr, ncam >= 2 ? F.mvRightToLeftMatch.data() : nullptr,
                                  stereo ? F.mvuRight.data() : nullptr, nmp, in_view.data(), px.data(), py.data(), level.data(), vc.data(),
                                  stereo ? pxr.data() : nullptr, depth.data(), bad.data(), has_obs.data(), mp_desc.data(), th, bFarPoints ? 1 : 0,
                                  thFarPoints, mfNNratio, assigned.data(), &nmatches) != ORBX_OK)
        die("SearchByProjection(Frame, MapPoints)");
    for (int i = 0; i < N; i++)
        if (assigned[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assigned[i]];          // :122, 127, 195-199, 267, 332
    return nmatches;
}

// src/ORBmatcher.cc:895-1004
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize) {
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f must be two packed floats");
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    vnMatches12 = std::vector<int>(n1, -1);                                           // :902
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<unsigned char> t1, t2;
    const unsigned char *d1 = desc_rows(F1.mDescriptors, t1), *d2 = desc_rows(F2.mDescriptors, t2);
    const float bounds[4] = {F2.mnMinX, F2.mnMaxX, F2.mnMinY, F2.mnMaxY};
    int nmatches = 0;
    if (orbx_search_for_initialization(handle(), (const orbx_kp *)F1.mvKeysUn.data(), d1, n1, (const orbx_kp *)F2.mvKeysUn.data(), d2, n2, bounds,
                                       (float *)vbPrevMatched.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, vnMatches12.data(),
                                       &nmatches) != ORBX_OK)
        die("SearchForInitialization");
    return nmatches;
}

// src/ORBmatcher.cc:349-666
int ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches) {
    const std::vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint *>(F.N, static_cast<MapPoint *>(NULL));
    const int nKF = (int)vpMapPointsKF.size(), nF = F.N;
    if (nKF == 0 || nF == 0) return 0;
    std::vector<int> nk, pk, ik, nf, pf, iff;
    feature_vector_csr(pKF->mFeatVec, nk, pk, ik);
    feature_vector_csr(F.mFeatVec, nf, pf, iff);
    std::vector<unsigned char> skip(nKF, 0), t1, t2;
    for (int i = 0; i < nKF; i++) skip[i] = !vpMapPointsKF[i] || vpMapPointsKF[i]->isBad();     // :377-380
    // keypoint angles by feature index, chosen as :484-505 choose kp / Fkp
    std::vector<float> aKF(nKF, 0.f), aF(nF, 0.f);
    if (mbCheckOrientation) {
        for (int i = 0; i < nKF; i++) {
            const cv::KeyPoint &kp = (!pKF->mpCamera2)                              ? pKF->mvKeysUn[i]
                                     : (i < pKF->NLeft)                            ? pKF->mvKeys[i]
                                     : (i < pKF->NLeft + pKF->NRight)              ? pKF->mvKeysRight[i - pKF->NLeft]
                                     : (i < pKF->NLeft + pKF->NRight + pKF->NSideLeft) ? pKF->mvKeysSideLeft[i - pKF->NLeft - pKF->NRight]
                                                                                   : pKF->mvKeysSideRight[i - pKF->NLeft - pKF->NRight - pKF->NSideLeft];
            aKF[i] = kp.angle;
        }
        for (int i = 0; i < nF; i++) {
            const cv::KeyPoint &kp = (!pKF->mpCamera2 || i < F.Nleft)              ? F.mvKeys[i]
                                     : (i < F.Nleft + F.Nright)                    ? F.mvKeysRight[i - F.Nleft]
                                     : (i < F.Nleft + F.Nright + F.Nsideleft)      ? F.mvKeysSideLeft[i - F.Nleft - F.Nright]
                                                                                   : F.mvKeysSideRight[i - F.Nleft - F.Nright - F.Nsideleft];
            aF[i] = kp.angle;
        }
    }
    int cam_off[5] = {0, nF, nF, nF, nF};
    int ncam = 1;
    if (F.Nleft != -1) {
        cam_off[1] = F.Nleft; cam_off[2] = F.Nleft + F.Nright;
        ncam = 2;
        if (F.Nsideleft != -1 || F.Nsideright != -1) { cam_off[3] = cam_off[2] + F.Nsideleft; cam_off[4] = cam_off[3] + F.Nsideright; ncam = 4; }
        else cam_off[3] = cam_off[4] = cam_off[2];
    }
    std::vector<int> matchF(nF, -1);
    int nmatches = 0;
    if (orbx_search_by_bow_frame(handle(), (int)nk.size(), nk.data(), pk.data(), ik.data(), (int)nf.size(), nf.data(), pf.data(), iff.data(),
                                 desc_rows(pKF->mDescriptors, t1), nKF, desc_rows(F.mDescriptors, t2), nF, skip.data(), aKF.data(), aF.data(), ncam,
                                 cam_off, mfNNratio, mbCheckOrientation ? 1 : 0, matchF.data(), &nmatches) != ORBX_OK)
        die("SearchByBoW(KeyFrame, Frame)");
    for (int i = 0; i < nF; i++)
        if (matchF[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[matchF[i]];
    return nmatches;
}

// src/ORBmatcher.cc:1006-1129
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12) {
    const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
    vpMatches12 = std::vector<MapPoint *>(n1, static_cast<MapPoint *>(NULL));          // :1017
    if (n1 == 0 || n2 == 0) return 0;
    std::vector<int> nk1, pk1, ik1, nk2, pk2, ik2;
    feature_vector_csr(pKF1->mFeatVec, nk1, pk1, ik1);
    feature_vector_csr(pKF2->mFeatVec, nk2, pk2, ik2);
    std::vector<unsigned char> skip1(n1), skip2(n2), t1, t2;
    for (int i = 0; i < n1; i++) skip1[i] = (pKF1->NLeft != -1 && i >= pKF1->N) || !vpMapPoints1[i] || vpMapPoints1[i]->isBad();   // :1036-1044
    for (int i = 0; i < n2; i++) skip2[i] = (pKF2->NLeft != -1 && i >= pKF2->N) || !vpMapPoints2[i] || vpMapPoints2[i]->isBad();   // :1052-1061
    std::vector<float> a1(n1, 0.f), a2(n2, 0.f);
    if (mbCheckOrientation) {                                                           // vKeysUn1[idx1].angle - vKeysUn2[bestIdx2].angle, :1087
        for (int i = 0; i < n1 && i < (int)pKF1->mvKeysUn.size(); i++) a1[i] = pKF1->mvKeysUn[i].angle;
        for (int i = 0; i < n2 && i < (int)pKF2->mvKeysUn.size(); i++) a2[i] = pKF2->mvKeysUn[i].angle;
    }
    if (pKF1->mDescriptors.rows < n1 || pKF2->mDescriptors.rows < n2) { fprintf(stderr, "ORBmatcher::SearchByBoW: fewer descriptor rows than map point slots\n"); abort(); }
    std::vector<int> match12(n1, -1);
    int nmatches = 0;
    if (orbx_search_by_bow_keyframes(handle(), (int)nk1.size(), nk1.data(), pk1.data(), ik1.data(), (int)nk2.size(), nk2.data(), pk2.data(), ik2.data(),
                                     desc_rows(pKF1->mDescriptors, t1), n1, desc_rows(pKF2->mDescriptors, t2), n2, skip1.data(), skip2.data(), a1.data(),
                                     a2.data(), mfNNratio, mbCheckOrientation ? 1 : 0, match12.data(), &nmatches) != ORBX_OK)
        die("SearchByBoW(KeyFrame, KeyFrame)");
    for (int i = 0; i < n1; i++)
        if (match12[i] >= 0) vpMatches12[i] = vpMapPoints2[match12[i]];                 // :1084
    return nmatches;
}

namespace {
// keypoint i of a Frame in global feature order (the reference's kpLF / nLastOctave selection, :2027-2038, 2107-2117)
const cv::KeyPoint &frame_key(const Frame &F, int i, bool undistorted_if_single) {
    if (F.Nleft == -1) return undistorted_if_single ? F.mvKeysUn[i] : F.mvKeys[i];
    if (i < F.Nleft) return F.mvKeys[i];
    if (i < F.Nleft + F.Nright) return F.mvKeysRight[i - F.Nleft];
    if (i < F.Nleft + F.Nright + F.Nsideleft) return F.mvKeysSideLeft[i - F.Nleft - F.Nright];
    return F.mvKeysSideRight[i - F.Nleft - F.Nright - F.Nsideleft];
}
}  // namespace

// src/ORBmatcher.cc:1985-2413 — the tracking matcher (Tracking::TrackWithMotionModel)
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono) {
    std::vector<orbx_kp> kps;
    int cam_off[5];
    const int ncam = frame_layout(CurrentFrame, kps, cam_off);
    const int N = CurrentFrame.N, npts = LastFrame.N;
    if (N == 0 || npts == 0) return 0;
    // :1998-2004
    const Sophus::SE3f Tcw = CurrentFrame.GetPose();
    const Eigen::Vector3f twc = Tcw.inverse().translation();
    const Sophus::SE3f Tlw = LastFrame.GetPose();
    const Eigen::Vector3f tlc = Tlw * twc;
    const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
    std::vector<unsigned char> valid(npts, 0), has_obs(npts, 0), pt_desc((size_t)npts * 32, 0);
    std::vector<float> px((size_t)npts * 4, 0.f), py((size_t)npts * 4, 0.f), pur(npts, 0.f), last_angle(npts, 0.f);
    std::vector<int> last_octave(npts, 0);
    for (int i = 0; i < npts; i++) {
        MapPoint *pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;                                  // :2007-2009
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const float invzc = 1.0 / x3Dc(2);                                              // :2016
        if (invzc < 0) continue;
        Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        valid[i] = 1;
        px[4 * i] = uv(0); py[4 * i] = uv(1);
        pur[i] = uv(0) - CurrentFrame.mbf * invzc;                                      // :2080
        if (CurrentFrame.Nleft != -1) {                                                 // :2132-2134 (projected with mpCamera, as the reference does)
            Eigen::Vector3f x3Dr = CurrentFrame.GetRelativePoseTrl() * x3Dc;
            Eigen::Vector2f uvr = CurrentFrame.mpCamera->project(x3Dr);
            px[4 * i + 1] = uvr(0); py[4 * i + 1] = uvr(1);
        }
        px[4 * i + 2] = px[4 * i + 3] = uv(0); py[4 * i + 2] = py[4 * i + 3] = uv(1);   // the side cameras are searched around the LEFT projection (:2250-2271)
        last_octave[i] = frame_key(LastFrame, i, false).octave;                         // :2027-2038 (mvKeys for a single-camera frame)
        last_angle[i] = frame_key(LastFrame, i, true).angle;                            // :2107-2117 (mvKeysUn for a single-camera frame)
        has_obs[i] = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&pt_desc[(size_t)i * 32], d.ptr<unsigned char>(0), 32);
    }
    std::vector<unsigned char> occupied(N, 0), dtmp;
    for (int i = 0; i < N; i++) occupied[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;   // :2072-2074
    float bounds[16];
    for (int c = 0; c < 4; c++) { bounds[4 * c] = CurrentFrame.mnMinX; bounds[4 * c + 1] = CurrentFrame.mnMaxX; bounds[4 * c + 2] = CurrentFrame.mnMinY; bounds[4 * c + 3] = CurrentFrame.mnMaxY; }
    const bool stereo = ncam == 1 && (int)CurrentFrame.mvuRight.size() == N;            // :2076
    std::vector<int> assigned(N, -1);
    int nmatches = 0;
    orbx_matcher *h = handle();
    if (orbx_search_by_projection_last(h, ncam, cam_off, kps.data(), desc_rows(CurrentFrame.mDescriptors, dtmp), bounds, CurrentFrame.mvScaleFactors.data(),
                                       (int)CurrentFrame.mvScaleFactors.size(), occupied.data(), stereo ? CurrentFrame.mvuRight.data() : nullptr, npts,
                                       valid.data(), px.data(), py.data(), pur.data(), last_octave.data(), last_angle.data(), has_obs.data(), pt_desc.data(),
                                       th, bForward ? ORBX_GATE_FORWARD : bBackward ? ORBX_GATE_BACKWARD : ORBX_GATE_AROUND, mbCheckOrientation ? 1 : 0,
                                       TH_HIGH, assigned.data(), &nmatches) != ORBX_OK)
        die("SearchByProjection(CurrentFrame, LastFrame)");
    std::vector<unsigned char> cleared(N, 0);
    if (mbCheckOrientation && orbx_search_by_projection_last_cleared(h, cleared.data(), N) != ORBX_OK) die("SearchByProjection(CurrentFrame, LastFrame)");
    for (int i = 0; i < N; i++) {
        if (assigned[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[assigned[i]];   // :2104
        else if (cleared[i]) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint *>(NULL);           // :2404
    }
    return nmatches;
}

// src/ORBmatcher.cc:2415-2535 — the relocalisation matcher (searches camera 0 only)
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th, const int ORBdist) {
    const Sophus::SE3f Tcw = CurrentFrame.GetPose();
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    const std::vector<MapPoint *> vpMPs = pKF->GetMapPointMatches();
    const int npts = (int)vpMPs.size();
    // camera 0 of the current frame: Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) searches mvKeysUn (Nleft == -1) or mvKeys[0, Nleft)
    const int n0 = CurrentFrame.Nleft == -1 ? CurrentFrame.N : CurrentFrame.Nleft;
    if (npts == 0 || n0 == 0) return 0;
    std::vector<orbx_kp> kps(n0);
    for (int i = 0; i < n0; i++) {
        const cv::KeyPoint &k = CurrentFrame.Nleft == -1 ? CurrentFrame.mvKeysUn[i] : CurrentFrame.mvKeys[i];
        memcpy(&kps[i], &k, sizeof(orbx_kp));
        if (i < (int)CurrentFrame.mvKeysUn.size()) kps[i].angle = CurrentFrame.mvKeysUn[i].angle;   // :2503
    }
    const int cam_off[2] = {0, n0};
    std::vector<unsigned char> valid(npts, 0), pt_desc((size_t)npts * 32, 0);
    std::vector<float> px((size_t)npts * 4, 0.f), py((size_t)npts * 4, 0.f), last_angle(npts, 0.f);   // [point][camera], camera 0 used
    std::vector<int> level(npts, 0);
    const int nlevels = (int)CurrentFrame.mvScaleFactors.size();
    for (int i = 0; i < npts; i++) {
        MapPoint *pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;                 // :2436-2437
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        Eigen::Vector3f PO = x3Dw - Ow;
        float dist3D = PO.norm();
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance) continue;                     // :2458
        const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
        if (nPredictedLevel < 0 || nPredictedLevel >= nlevels) continue;                // (the reference would index mvScaleFactors out of range)
        valid[i] = 1; px[4 * i] = uv(0); py[4 * i] = uv(1); level[i] = nPredictedLevel;
        last_angle[i] = i < (int)pKF->mvKeysUn.size() ? pKF->mvKeysUn[i].angle : 0.f;   // :2503
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&pt_desc[(size_t)i * 32], d.ptr<unsigned char>(0), 32);
    }
    std::vector<unsigned char> occupied(n0, 0), dtmp;
    for (int i = 0; i < n0; i++) occupied[i] = CurrentFrame.mvpMapPoints[i] != NULL;    // :2484
    const float bounds[4] = {CurrentFrame.mnMinX, CurrentFrame.mnMaxX, CurrentFrame.mnMinY, CurrentFrame.mnMaxY};
    std::vector<int> assigned(n0, -1);
    int nmatches = 0;
    if (orbx_search_by_projection_last(handle(), 1, cam_off, kps.data(), desc_rows(CurrentFrame.mDescriptors, dtmp), bounds, CurrentFrame.mvScaleFactors.data(),
                                       nlevels, occupied.data(), nullptr, npts, valid.data(), px.data(), py.data(), nullptr, level.data(), last_angle.data(),
                                       nullptr, pt_desc.data(), th, ORBX_GATE_AROUND, mbCheckOrientation ? 1 : 0, ORBdist, assigned.data(), &nmatches) != ORBX_OK)
        die("SearchByProjection(CurrentFrame, KeyFrame)");
    // (only keypoints that held NULL can be claimed, so a claim cleared by the rotation filter, :2527, leaves what was there: NULL)
    for (int i = 0; i < n0; i++)
        if (assigned[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[assigned[i]];       // :2499
    return nmatches;
}

namespace {
// The loop-closing matchers (:668-777 and :779-893 differ only in vpMatchedKF): candidate map points projected into ONE camera of a
// KeyFrame under a Sim3, each matched to the most similar still-unmatched keypoint of that camera's block within [level - 1, level].
int search_keyframe_sim3(KeyFrame *pKF, Sophus::Sim3f &Siw, const std::vector<MapPoint *> &vpPoints, const std::vector<KeyFrame *> *vpPointsKFs,
                         std::vector<MapPoint *> &vpMatched, std::vector<KeyFrame *> *vpMatchedKF, int th, float ratioHamming, int cameraID) {
    Sophus::SE3f Tiw = Sophus::SE3f(Siw.rotationMatrix(), Siw.translation() / Siw.scale());   // :682
    Eigen::Vector3f Ow = Tiw.inverse().translation();
    std::set<MapPoint *> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint *>(NULL));
    // the camera's block of the KeyFrame (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:771-834: global index = block offset + grid index)
    const std::vector<cv::KeyPoint> &keys = pKF->NLeft == -1 ? pKF->mvKeysUn : cameraID == 0 ? pKF->mvKeys : cameraID == 1 ? pKF->mvKeysRight
                                            : cameraID == 2 ? pKF->mvKeysSideLeft : pKF->mvKeysSideRight;
    const int off = pKF->NLeft == -1 || cameraID == 0 ? 0 : cameraID == 1 ? pKF->NLeft : cameraID == 2 ? pKF->NLeft + pKF->NRight
                                                                                     : pKF->NLeft + pKF->NRight + pKF->NSideLeft;
    const int n = pKF->NLeft == -1 ? pKF->N : (int)keys.size();
    const int npts = (int)vpPoints.size();
    if (n == 0 || npts == 0) return 0;
    const int nlevels = (int)pKF->mvScaleFactors.size();
    std::vector<unsigned char> valid(npts, 0), pt_desc((size_t)npts * 32, 0);
    std::vector<float> px((size_t)npts * 4, 0.f), py((size_t)npts * 4, 0.f), angle(npts, 0.f);
    std::vector<int> level(npts, 0);
    for (int iMP = 0; iMP < npts; iMP++) {
        MapPoint *pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;                        // :696
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tiw * p3Dw;
        if (p3Dc(2) < 0.0) continue;                                                    // :706
        const Eigen::Vector2f uv = pKF->GetCamera(cameraID)->project(p3Dc);
        if (!pKF->IsInImage(uv(0), uv(1))) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist = PO.norm();
        if (dist < minDistance || dist > maxDistance) continue;
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;                                          // :728
        const int nPredictedLevel = pMP->PredictScale(dist, pKF);
        if (nPredictedLevel < 0 || nPredictedLevel >= nlevels) continue;                // (the reference would index mvScaleFactors out of range)
        valid[iMP] = 1; px[4 * iMP] = uv(0); py[4 * iMP] = uv(1); level[iMP] = nPredictedLevel;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&pt_desc[(size_t)iMP * 32], d.ptr<unsigned char>(0), 32);
    }
    std::vector<unsigned char> occupied(n, 0), dtmp;
    for (int i = 0; i < n; i++) occupied[i] = vpMatched[off + i] != NULL;               // :749
    const unsigned char *desc = desc_rows(pKF->mDescriptors, dtmp) + (size_t)off * 32;
    const int cam_off[2] = {0, n};
    const float bounds[4] = {(float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY};
    const int max_dist = (int)floorf(ORBmatcher::TH_LOW * ratioHamming);                // bestDist <= TH_LOW * ratioHamming, :767
    std::vector<int> assigned(n, -1);
    int nmatches = 0;
    if (orbx_search_by_projection_last(handle(), 1, cam_off, (const orbx_kp *)keys.data(), desc, bounds, pKF->mvScaleFactors.data(), nlevels, occupied.data(),
                                       nullptr, npts, valid.data(), px.data(), py.data(), nullptr, level.data(), angle.data(), nullptr, pt_desc.data(),
                                       (float)th, ORBX_GATE_BELOW, 0, max_dist < 0 ? 0 : max_dist > 255 ? 255 : max_dist, assigned.data(), &nmatches) != ORBX_OK)
        die("SearchByProjection(KeyFrame, Sim3)");
    for (int i = 0; i < n; i++)
        if (assigned[i] >= 0) {
            vpMatched[off + i] = vpPoints[assigned[i]];                                 // :768
            if (vpMatchedKF) (*vpMatchedKF)[off + i] = (*vpPointsKFs)[assigned[i]];     // :888
        }
    return nmatches;
}
}  // namespace

// src/ORBmatcher.cc:668-777
int ORBmatcher::SearchByProjection(KeyFrame *pKF, Sophus::Sim3f &Siw, const std::vector<MapPoint *> &vpPoints, std::vector<MapPoint *> &vpMatched, int th,
                                   float ratioHamming, int cameraID) {
    return search_keyframe_sim3(pKF, Siw, vpPoints, nullptr, vpMatched, nullptr, th, ratioHamming, cameraID);
}

// src/ORBmatcher.cc:779-893
int ORBmatcher::SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Siw, const std::vector<MapPoint *> &vpPoints, const std::vector<KeyFrame *> &vpPointsKFs,
                                   std::vector<MapPoint *> &vpMatched, std::vector<KeyFrame *> &vpMatchedKF, int th, float ratioHamming, int cameraID) {
    return search_keyframe_sim3(pKF, Siw, vpPoints, &vpPointsKFs, vpMatched, &vpMatchedKF, th, ratioHamming, cameraID);
}

// src/ORBmatcher.cc:1458-1647 — LocalMapping::SearchInNeighbors.  The searches of all points run in one device call (a search reads
// keypoints and descriptors only, nothing the graph updates change); the updates follow in point order with the reference's
// dynamic tests (isBad / IsInKeyFrame / GetMapPoint) evaluated at that moment.
int ORBmatcher::Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th, const int cameraID) {
    GeometricCamera *pCamera = nullptr;
    Sophus::SE3f Tcw;
    Eigen::Vector3f Ow;
    switch (cameraID) {                                                                  // :1467-1495
    case 0: Tcw = pKF->GetPose(); Ow = pKF->GetCameraCenter(); pCamera = pKF->mpCamera; break;
    case 1: Tcw = pKF->GetRightPose(); Ow = pKF->GetRightCameraCenter(); pCamera = pKF->mpCamera2; break;
    case 2: Tcw = pKF->GetSideLeftPose(); Ow = pKF->GetSideLeftCameraCenter(); pCamera = pKF->mpCamera3; break;
    case 3: Tcw = pKF->GetSideRightPose(); Ow = pKF->GetSideRightCameraCenter(); pCamera = pKF->mpCamera4; break;
    default: break;
    }
    const float &bf = pKF->mbf;
    const std::vector<cv::KeyPoint> &keys = pKF->NLeft == -1 ? pKF->mvKeysUn : cameraID == 0 ? pKF->mvKeys : cameraID == 1 ? pKF->mvKeysRight
                                            : cameraID == 2 ? pKF->mvKeysSideLeft : pKF->mvKeysSideRight;
    const int off = pKF->NLeft == -1 || cameraID == 0 ? 0 : cameraID == 1 ? pKF->NLeft : cameraID == 2 ? pKF->NLeft + pKF->NRight
                                                                                     : pKF->NLeft + pKF->NRight + pKF->NSideLeft;
    const int n = pKF->NLeft == -1 ? pKF->N : (int)keys.size();
    const int nMPs = (int)vpMapPoints.size();
    if (!pCamera || n == 0 || nMPs == 0) return 0;
    const int nlevels = (int)pKF->mvScaleFactors.size();
    std::vector<unsigned char> valid(nMPs, 0), q_desc((size_t)nMPs * 32, 0);
    std::vector<float> qx(nMPs, 0.f), qy(nMPs, 0.f), qur(nMPs, 0.f), radius(nMPs, 0.f);
    std::vector<int> min_level(nMPs, 0), max_level(nMPs, 0);
    for (int i = 0; i < nMPs; i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                    // (tested again, at its turn, below)
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tcw * p3Dw;
        if (p3Dc(2) < 0.0f) continue;                                                    // :1528
        const float invz = 1 / p3Dc(2);
        const Eigen::Vector2f uv = pCamera->project(p3Dc);
        if (!pKF->IsInImage(uv(0), uv(1))) continue;
        const float ur = uv(0) - bf * invz;                                              // :1543
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist3D = PO.norm();
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) continue;                                         // :1560
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        if (nPredictedLevel < 0 || nPredictedLevel >= nlevels) continue;                 // (the reference would index mvScaleFactors out of range)
        valid[i] = 1; qx[i] = uv(0); qy[i] = uv(1); qur[i] = ur;
        radius[i] = th * pKF->mvScaleFactors[nPredictedLevel];                           // :1568
        min_level[i] = nPredictedLevel - 1; max_level[i] = nPredictedLevel;              // :1592
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&q_desc[(size_t)i * 32], d.ptr<unsigned char>(0), 32);
    }
    // the stereo reprojection test applies to camera 0 only (:1595): other cameras pass no right coordinates
    const bool stereo = cameraID == 0 && (int)pKF->mvuRight.size() >= off + n;
    std::vector<unsigned char> dtmp;
    const float bounds[4] = {(float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY};
    std::vector<int> best_idx(nMPs, -1), best_dist(nMPs, 256);
    if (orbx_window_best(handle(), (const orbx_kp *)keys.data(), desc_rows(pKF->mDescriptors, dtmp) + (size_t)off * 32, n, bounds, nullptr,
                         stereo ? pKF->mvuRight.data() + off : nullptr, pKF->mvInvLevelSigma2.data(), nlevels, nMPs, valid.data(), qx.data(), qy.data(),
                         qur.data(), radius.data(), min_level.data(), max_level.data(), q_desc.data(), best_idx.data(), best_dist.data()) != ORBX_OK)
        die("Fuse(KeyFrame, MapPoints)");
    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {
        if (!valid[i]) continue;
        MapPoint *pMP = vpMapPoints[i];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                            // :1511-1517, now that the earlier points have been fused
        if (best_idx[i] < 0 || best_dist[i] > TH_LOW) continue;                          // :1628
        const int bestIdx = off + best_idx[i];
        MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

namespace {
// one camera's block of a KeyFrame: its keypoints, the offset of its first feature in the global order, its size
const std::vector<cv::KeyPoint> &keyframe_block(KeyFrame *pKF, int cameraID, int &off, int &n) {
    const std::vector<cv::KeyPoint> &keys = pKF->NLeft == -1 ? pKF->mvKeysUn : cameraID == 0 ? pKF->mvKeys : cameraID == 1 ? pKF->mvKeysRight
                                            : cameraID == 2 ? pKF->mvKeysSideLeft : pKF->mvKeysSideRight;
    off = pKF->NLeft == -1 || cameraID == 0 ? 0 : cameraID == 1 ? pKF->NLeft : cameraID == 2 ? pKF->NLeft + pKF->NRight : pKF->NLeft + pKF->NRight + pKF->NSideLeft;
    n = pKF->NLeft == -1 ? (cameraID == 0 ? pKF->N : 0) : (int)keys.size();     // (a single-camera keyframe has empty grids for cameras 1-3)
    return keys;
}
}  // namespace

// src/ORBmatcher.cc:1648-1769 — the loop-closing fuse: every candidate point is searched in all four cameras of the keyframe
int ORBmatcher::Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, const std::vector<MapPoint *> &vpPoints, float th, std::vector<MapPoint *> &vpReplacePoint) {
    const std::set<MapPoint *> spAlreadyFound = pKF->GetMapPoints();
    const int nPoints = (int)vpPoints.size();
    if (nPoints == 0) return 0;
    const int nlevels = (int)pKF->mvScaleFactors.size();
    std::vector<unsigned char> usable(nPoints, 0), q_desc((size_t)nPoints * 32, 0), dtmp;
    for (int i = 0; i < nPoints; i++) {
        MapPoint *pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;                         // :1675
        usable[i] = 1;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&q_desc[(size_t)i * 32], d.ptr<unsigned char>(0), 32);
    }
    const unsigned char *desc = desc_rows(pKF->mDescriptors, dtmp);
    const float bounds[4] = {(float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY};
    std::vector<unsigned char> valid[4];
    std::vector<int> best_idx[4], best_dist[4];
    int off[4] = {0, 0, 0, 0};
    Sophus::SE3f Tlw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());   // :1693
    for (int camId = 0; camId < 4; camId++) {
        int n;
        const std::vector<cv::KeyPoint> &keys = keyframe_block(pKF, camId, off[camId], n);
        valid[camId].assign(nPoints, 0); best_idx[camId].assign(nPoints, -1); best_dist[camId].assign(nPoints, 256);
        if (n == 0) continue;
        Sophus::SE3f Tiw = (camId == 0) ? Tlw : (camId == 1) ? pKF->GetRelativePoseTrl() * Tlw : (camId == 2) ? pKF->GetRelativePoseTsll() * Tlw
                                                                                               : pKF->GetRelativePoseTsrl() * Tlw;
        Eigen::Vector3f Ow = Tiw.inverse().translation();
        std::vector<float> qx(nPoints, 0.f), qy(nPoints, 0.f), radius(nPoints, 0.f);
        std::vector<int> min_level(nPoints, 0), max_level(nPoints, 0);
        for (int i = 0; i < nPoints; i++) {
            if (!usable[i]) continue;
            MapPoint *pMP = vpPoints[i];
            Eigen::Vector3f p3Dw = pMP->GetWorldPos();
            Eigen::Vector3f p3Dc = Tiw * p3Dw;
            if (p3Dc(2) < 0.0f) continue;                                                // :1707
            const Eigen::Vector2f uv = pKF->GetCamera(camId)->project(p3Dc);
            if (!pKF->IsInImage(uv(0), uv(1))) continue;
            Eigen::Vector3f PO = p3Dw - Ow;
            const float dist3D = PO.norm();
            const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);                  // (no distance / viewing-angle gate in this overload)
            if (nPredictedLevel < 0 || nPredictedLevel >= nlevels) continue;
            valid[camId][i] = 1; qx[i] = uv(0); qy[i] = uv(1);
            radius[i] = th * pKF->mvScaleFactors[nPredictedLevel];
            min_level[i] = nPredictedLevel - 1; max_level[i] = nPredictedLevel;
        }
        if (orbx_window_best(handle(), (const orbx_kp *)keys.data(), desc + (size_t)off[camId] * 32, n, bounds, nullptr, nullptr, nullptr, nlevels, nPoints,
                             valid[camId].data(), qx.data(), qy.data(), nullptr, radius.data(), min_level.data(), max_level.data(), q_desc.data(),
                             best_idx[camId].data(), best_dist[camId].data()) != ORBX_OK)
            die("Fuse(KeyFrame, Sim3)");
    }
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {
        if (!usable[iMP]) continue;
        MapPoint *pMP = vpPoints[iMP];
        for (int camId = 0; camId < 4; camId++) {
            if (!valid[camId][iMP] || best_idx[camId][iMP] < 0 || best_dist[camId][iMP] > TH_LOW) continue;   // :1750
            const int bestIdx = off[camId] + best_idx[camId][iMP];
            MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
    }
    return nFused;
}

// src/ORBmatcher.cc:1771-1983
int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, const Sophus::Sim3f &S12, const float th) {
    const float &fx = pKF1->fx;
    const float &fy = pKF1->fy;
    const float &cx = pKF1->cx;
    const float &cy = pKF1->cy;
    Sophus::SE3f T1w = pKF1->GetPose();
    Sophus::SE3f T2w = pKF2->GetPose();
    Sophus::Sim3f S21 = S12.inverse();
    const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size();
    const std::vector<MapPoint *> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = (int)vpMapPoints2.size();
    if (N1 == 0 || N2 == 0) return 0;
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {                                                       // :1798-1806
        MapPoint *pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    // one direction: the points of `from` (pose Tfw) mapped by S into the camera of `to`
    auto side = [&](const std::vector<MapPoint *> &pts, const std::vector<bool> &already, const Sophus::SE3f &Tfw, const Sophus::Sim3f &S, KeyFrame *to,
                    std::vector<unsigned char> &valid, std::vector<float> &u, std::vector<float> &v, std::vector<float> &radius, std::vector<int> &level,
                    std::vector<unsigned char> &mp_desc) {
        const int n = (int)pts.size(), nlevels = (int)to->mvScaleFactors.size();
        valid.assign(n, 0); u.assign(n, 0.f); v.assign(n, 0.f); radius.assign(n, 0.f); level.assign(n, 0); mp_desc.assign((size_t)n * 32, 0);
        for (int i = 0; i < n; i++) {
            MapPoint *pMP = pts[i];
            if (!pMP || already[i]) continue;
            if (pMP->isBad()) continue;
            Eigen::Vector3f p3Dw = pMP->GetWorldPos();
            Eigen::Vector3f p3Dcf = Tfw * p3Dw;
            Eigen::Vector3f p3Dct = S * p3Dcf;
            if (p3Dct(2) < 0.0) continue;                                                // :1827
            const float invz = 1.0 / p3Dct(2);
            const float x = p3Dct(0) * invz;
            const float y = p3Dct(1) * invz;
            const float uu = fx * x + cx;
            const float vv = fy * y + cy;
            if (!to->IsInImage(uu, vv)) continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            const float dist3D = p3Dct.norm();
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, to);
            if (nPredictedLevel < 0 || nPredictedLevel >= nlevels) continue;
            valid[i] = 1; u[i] = uu; v[i] = vv; level[i] = nPredictedLevel;
            radius[i] = th * to->mvScaleFactors[nPredictedLevel];
            const cv::Mat d = pMP->GetDescriptor();
            memcpy(&mp_desc[(size_t)i * 32], d.ptr<unsigned char>(0), 32);
        }
    };
    std::vector<unsigned char> valid1, valid2, md1, md2, t1, t2;
    std::vector<float> u1, v1, r1, u2, v2, r2;
    std::vector<int> l1, l2;
    side(vpMapPoints1, vbAlreadyMatched1, T1w, S21, pKF2, valid1, u1, v1, r1, l1, md1);  // KF1's points searched in KF2 (:1812-1887)
    side(vpMapPoints2, vbAlreadyMatched2, T2w, S12, pKF1, valid2, u2, v2, r2, l2, md2);  // KF2's points searched in KF1 (:1890-1965)
    const float b1[4] = {(float)pKF1->mnMinX, (float)pKF1->mnMaxX, (float)pKF1->mnMinY, (float)pKF1->mnMaxY};
    const float b2[4] = {(float)pKF2->mnMinX, (float)pKF2->mnMaxX, (float)pKF2->mnMinY, (float)pKF2->mnMaxY};
    if ((int)pKF1->mvKeysUn.size() < N1 || (int)pKF2->mvKeysUn.size() < N2) { fprintf(stderr, "ORBmatcher::SearchBySim3: fewer undistorted keypoints than map point slots\n"); abort(); }
    std::vector<int> match12(N1, -1);
    int nFound = 0;
    if (orbx_search_by_sim3(handle(), (const orbx_kp *)pKF1->mvKeysUn.data(), desc_rows(pKF1->mDescriptors, t1), N1, b1, (const orbx_kp *)pKF2->mvKeysUn.data(),
                            desc_rows(pKF2->mDescriptors, t2), N2, b2, valid1.data(), u1.data(), v1.data(), r1.data(), l1.data(), md1.data(), valid2.data(),
                            u2.data(), v2.data(), r2.data(), l2.data(), md2.data(), match12.data(), &nFound) != ORBX_OK)
        die("SearchBySim3");
    for (int i1 = 0; i1 < N1; i1++)
        if (match12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[match12[i1]];              // :1975
    return nFound;
}

namespace {
// keypoints of a KeyFrame in global feature order and its camera blocks (as frame_layout does for a Frame)
int keyframe_layout(KeyFrame *K, std::vector<orbx_kp> &kps, int cam_off[5]) {
    auto add = [&](const std::vector<cv::KeyPoint> &v, int n) { kps.insert(kps.end(), (const orbx_kp *)v.data(), (const orbx_kp *)v.data() + n); };
    kps.clear();
    cam_off[0] = 0;
    if (K->NLeft == -1) {
        add(K->mvKeysUn, K->N);
        for (int c = 1; c <= 4; c++) cam_off[c] = K->N;
        return 1;
    }
    add(K->mvKeys, K->NLeft); cam_off[1] = K->NLeft;
    add(K->mvKeysRight, K->NRight); cam_off[2] = K->NLeft + K->NRight;
    if (K->NSideLeft == -1 && K->NSideRight == -1) { cam_off[3] = cam_off[4] = cam_off[2]; return 2; }
    add(K->mvKeysSideLeft, K->NSideLeft); cam_off[3] = cam_off[2] + K->NSideLeft;
    add(K->mvKeysSideRight, K->NSideRight); cam_off[4] = cam_off[3] + K->NSideRight;
    return 4;
}

// one camera pair of SearchForTriangulation as the device predicate takes it: Pinhole::epipolarConstrain's fundamental matrix
// (src/CameraModels/Pinhole.cpp:112-115, the reference's own expression, evaluated once per pair instead of once per candidate) or the
// Kannala-Brandt parameters of KannalaBrandt8::epipolarConstrain (src/CameraModels/KannalaBrandt8.cpp:219-230)
void epipolar_pair(orbx_epipolar_pair &P, GeometricCamera *c1, GeometricCamera *c2, const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12) {
    memset(&P, 0, sizeof(P));
    if (!c1 || !c2) return;
    if (c1->GetType() == GeometricCamera::CAM_PINHOLE) {
        P.model = ORBX_CAM_PINHOLE;
        Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
        Eigen::Matrix3f K1 = c1->toK_();
        Eigen::Matrix3f K2 = c2->toK_();
        Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P.F12[3 * i + j] = F12(i, j);
    } else {
        P.model = ORBX_CAM_KB8;
        for (int i = 0; i < 8; i++) { P.kb8.cam1[i] = c1->getParameter(i); P.kb8.cam2[i] = c2->getParameter(i); }
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) P.kb8.R12[3 * i + j] = R12(i, j); P.kb8.t12[i] = t12(i); }
#ifdef ORBX_GLUE_BUILD
        P.kb8.precision1 = P.kb8.precision2 = 1e-6f;           // (the test harness has no KannalaBrandt8 class; KannalaBrandt8.h:21)
#else
        P.kb8.precision1 = static_cast<KannalaBrandt8 *>(c1)->GetPrecision(); P.kb8.precision2 = static_cast<KannalaBrandt8 *>(c2)->GetPrecision();
#endif
    }
}
}  // namespace

// src/ORBmatcher.cc:1131-1456 — LocalMapping::CreateNewMapPoints
int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<std::pair<size_t, size_t>> &vMatchedPairs, const bool bOnlyStereo,
                                       const bool bCoarse) {
    // :1141-1198, the reference's expressions
    Sophus::SE3f T1w = pKF1->GetPose();
    Sophus::SE3f T2w = pKF2->GetPose();
    Sophus::SE3f Tw2 = pKF2->GetPoseInverse();
    Eigen::Vector3f Cw = pKF1->GetCameraCenter();
    Eigen::Vector3f C2 = T2w * Cw;
    Eigen::Vector2f ep = pKF2->mpCamera->project(C2);
    Sophus::SE3f T12, Tll, Tlr, Trl, Trr, Tlsl, Tsll, Tslsl, Trsr, Tsrr, Tsrsr;
    const bool rig2 = pKF1->mpCamera2 || pKF2->mpCamera2;                                // (:1156: the single-camera branch needs both to be NULL)
    const bool both2 = pKF1->mpCamera2 && pKF2->mpCamera2;
    const bool rig4 = pKF1->mpCamera3 && pKF2->mpCamera3 && pKF1->mpCamera4 && pKF2->mpCamera4;
    if (!rig2) {
        T12 = T1w * Tw2;
    } else {
        Sophus::SE3f Tr1w = pKF1->GetRightPose();
        Sophus::SE3f Twr2 = pKF2->GetRightPoseInverse();
        Tll = T1w * Tw2; Tlr = T1w * Twr2; Trl = Tr1w * Tw2; Trr = Tr1w * Twr2;
    }
    if (rig4) {
        Sophus::SE3f Tr1w = pKF1->GetRightPose();
        Sophus::SE3f Twr2 = pKF2->GetRightPoseInverse();
        Sophus::SE3f Tsl1w = pKF1->GetSideLeftPose();
        Sophus::SE3f Twsl2 = pKF2->GetSideLeftPoseInverse();
        Sophus::SE3f Tsr1w = pKF1->GetSideRightPose();
        Sophus::SE3f Twsr2 = pKF2->GetSideRightPoseInverse();
        Tlsl = T1w * Twsl2; Tsll = Tsl1w * Tw2; Tslsl = Tsl1w * Twsl2;
        Trsr = Tr1w * Twsr2; Tsrr = Tsr1w * Twr2; Tsrsr = Tsr1w * Twsr2;
    }
    std::vector<orbx_kp> kps1, kps2;
    int off1[5], off2[5];
    const int ncam1 = keyframe_layout(pKF1, kps1, off1), ncam2 = keyframe_layout(pKF2, kps2, off2);
    const int ncam = !both2 ? 1 : rig4 ? 4 : 2;                                          // the branches the reference's loop can take (:1301, 1331)
    if (ncam1 != ncam2 || ncam1 != ncam) { fprintf(stderr, "ORBmatcher::SearchForTriangulation: keyframes of different rigs are not supported\n"); abort(); }
    const int n1 = pKF1->N, n2 = pKF2->N;
    vMatchedPairs.clear();
    if (n1 == 0 || n2 == 0) return 0;
    // [0] the single-camera pair, [1] ll, [2] lr, [3] rl, [4] rr, [5] l-sl, [6] sl-l, [7] sl-sl, [8] r-sr, [9] sr-r, [10] sr-sr (:1301-1378)
    orbx_epipolar_pair pairs[11];
    memset(pairs, 0, sizeof(pairs));
    if (!bCoarse) {
        GeometricCamera *a0 = pKF1->mpCamera, *a1 = pKF1->mpCamera2, *a2 = pKF1->mpCamera3, *a3 = pKF1->mpCamera4;
        GeometricCamera *b0 = pKF2->mpCamera, *b1 = pKF2->mpCamera2, *b2 = pKF2->mpCamera3, *b3 = pKF2->mpCamera4;
        if (ncam == 1) epipolar_pair(pairs[0], a0, b0, T12.rotationMatrix(), T12.translation());
        if (ncam >= 2) {
            epipolar_pair(pairs[1], a0, b0, Tll.rotationMatrix(), Tll.translation());
            epipolar_pair(pairs[2], a0, b1, Tlr.rotationMatrix(), Tlr.translation());
            epipolar_pair(pairs[3], a1, b0, Trl.rotationMatrix(), Trl.translation());
            epipolar_pair(pairs[4], a1, b1, Trr.rotationMatrix(), Trr.translation());
        }
        if (ncam == 4) {
            epipolar_pair(pairs[5], a0, b2, Tlsl.rotationMatrix(), Tlsl.translation());
            epipolar_pair(pairs[6], a2, b0, Tsll.rotationMatrix(), Tsll.translation());
            epipolar_pair(pairs[7], a2, b2, Tslsl.rotationMatrix(), Tslsl.translation());
            epipolar_pair(pairs[8], a1, b3, Trsr.rotationMatrix(), Trsr.translation());
            epipolar_pair(pairs[9], a3, b1, Tsrr.rotationMatrix(), Tsrr.translation());
            epipolar_pair(pairs[10], a3, b3, Tsrsr.rotationMatrix(), Tsrsr.translation());
        }
    }
    std::vector<int> nk1, pk1, ik1, nk2, pk2, ik2;
    feature_vector_csr(pKF1->mFeatVec, nk1, pk1, ik1);
    feature_vector_csr(pKF2->mFeatVec, nk2, pk2, ik2);
    std::vector<unsigned char> mp1(n1), mp2(n2), t1, t2;
    for (int i = 0; i < n1; i++) mp1[i] = pKF1->GetMapPoint(i) != NULL;                  // :1224-1229
    for (int i = 0; i < n2; i++) mp2[i] = pKF2->GetMapPoint(i) != NULL;                  // :1260-1264
    const int nlevels = (int)pKF2->mvScaleFactors.size();
    std::vector<int> match12(n1, -1);
    int nmatches = 0;
    const float epf[2] = {ep(0), ep(1)};
    if (orbx_search_for_triangulation_full(handle(), ncam, off1, kps1.data(), desc_rows(pKF1->mDescriptors, t1), mp1.data(),
                                           ncam == 1 ? pKF1->mvuRight.data() : nullptr, (int)nk1.size(), nk1.data(), pk1.data(), ik1.data(), off2, kps2.data(),
                                           desc_rows(pKF2->mDescriptors, t2), mp2.data(), ncam == 1 ? pKF2->mvuRight.data() : nullptr, (int)nk2.size(),
                                           nk2.data(), pk2.data(), ik2.data(), epf, pKF2->mvScaleFactors.data(), pKF1->mvLevelSigma2.data(),
                                           pKF2->mvLevelSigma2.data(), nlevels, bOnlyStereo ? 1 : 0, bCoarse ? 1 : 0, mbCheckOrientation ? 1 : 0,
                                           bCoarse ? nullptr : pairs, match12.data(), &nmatches) != ORBX_OK)
        die("SearchForTriangulation");
    vMatchedPairs.reserve(nmatches);                                                     // :1446-1453
    for (int i = 0; i < n1; i++)
        if (match12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)match12[i]));
    return nmatches;
}

}  // namespace ORB_SLAM3
