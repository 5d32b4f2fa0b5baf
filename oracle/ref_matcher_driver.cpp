// oracle/ref_matcher_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the UNMODIFIED reference class
// ORB_SLAM3::ORBmatcher (oracle/Makefile compiles /root/reference/src/ORBmatcher.cc against oracle/refstub + oracle/cvshim) and the
// grid functions of Frame / KeyFrame, whose bodies the Makefile lifts from src/Frame.cc / src/KeyFrame.cc at build time into
// oracle/_ref/gen/*.inc (included below; never stored in the repo).  Every entry point takes the same flat arrays as the matching
// orc_* function of oracle/orb_oracle.cpp, builds stub Frame / KeyFrame / MapPoint objects from them and calls the reference's own
// member function; tests compare reference == oracle == CUDA on the same inputs.
// Geometry is injected (oracle/refstub/orb_slam3_stubs.h): identity poses and a "pixel camera" (project(X) = (X.x, X.y)), so the
// projections the reference computes are exactly the numbers the array-level interfaces take as inputs.
#include <stdint.h>
#include <string.h>

#include <memory>
#include <set>
#include <vector>

#include "ORBmatcher.h"

namespace ORB_SLAM3 {
float Frame::fx = 1, Frame::fy = 1, Frame::cx = 0, Frame::cy = 0, Frame::invfx = 1, Frame::invfy = 1;
float Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0;
// reference bodies, lifted at build time (oracle/Makefile: sed -n '<range>p' of the mounted sources)
#include "_ref/gen/frame_grid.inc"      // src/Frame.cc:541-582 AssignFeaturesToGrid, :890-967 GetFeaturesInArea, :969-978 PosInGrid
#include "_ref/gen/keyframe_grid.inc"   // src/KeyFrame.cc:769-834 GetFeaturesInArea

// A pinhole camera whose epipolarConstrain is the REFERENCE's (src/CameraModels/Pinhole.cpp:103-132, lifted at build time): project() and
// toK_() are the two obvious lines of include/CameraModels/Pinhole.h / Pinhole.cpp:28-35, 95-100 on the parameters {fx, fy, cx, cy}.
class Pinhole : public GeometricCamera {
  public:
    Pinhole(float fx, float fy, float cx, float cy) { mvParameters = {fx, fy, cx, cy}; mnType = CAM_PINHOLE; }
    Eigen::Vector2f project(const Eigen::Vector3f &v3D) override {
        return Eigen::Vector2f(mvParameters[0] * v3D[0] / v3D[2] + mvParameters[2], mvParameters[1] * v3D[1] / v3D[2] + mvParameters[3]);
    }
    Eigen::Matrix3f toK_() override {
        Eigen::Matrix3f K;
        K(0, 0) = mvParameters[0]; K(0, 2) = mvParameters[2]; K(1, 1) = mvParameters[1]; K(1, 2) = mvParameters[3]; K(2, 2) = 1.f;
        return K;
    }
    bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                           const Eigen::Vector3f &t12, const float sigmaLevel, const float unc) override;
};
#include "_ref/gen/pinhole_epipolar.inc"   // src/CameraModels/Pinhole.cpp:103-132 Pinhole::epipolarConstrain
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {

struct ref_kp { float x, y, size, angle, response; int octave, class_id; };

// project(X) = (X.x, X.y); epipolarConstrain = a deterministic hash of the two keypoints and the two camera ids, so that a test sees
// WHICH camera pair the reference's carried (pCamera1, pCamera2, R12, t12) state held when it evaluated a candidate
// (src/ORBmatcher.cc:1301-1390)
struct PixelCamera : GeometricCamera {
    int id;
    explicit PixelCamera(int i) : id(i) {}
    Eigen::Vector2f project(const Eigen::Vector3f &v) override { return Eigen::Vector2f(v(0), v(1)); }
    bool epipolarConstrain(GeometricCamera *c2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &, const Eigen::Vector3f &,
                           const float, const float) override {
        return pair_hash(kp1.pt.x, kp1.pt.y, kp2.pt.x, kp2.pt.y, id, static_cast<PixelCamera *>(c2)->id);
    }
    static bool pair_hash(float x1, float y1, float x2, float y2, int id1, int id2) {
        uint32_t a, b, c, d;
        memcpy(&a, &x1, 4); memcpy(&b, &y1, 4); memcpy(&c, &x2, 4); memcpy(&d, &y2, 4);
        uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u ^ c * 0xC2B2AE3Du ^ d * 0x27D4EB2Fu ^ (uint32_t)(id1 * 4 + id2) * 0x165667B1u;
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
        return (h & 3u) != 0;       // three out of four pairs pass
    }
};

cv::KeyPoint to_cv(const ref_kp &k) { return cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id); }

cv::Mat desc_mat(const uint8_t *d, int n) {
    cv::Mat m(n, 32, CV_8U);
    if (n) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

void set_bounds(const float *b) {
    Frame::mnMinX = b[0]; Frame::mnMaxX = b[1]; Frame::mnMinY = b[2]; Frame::mnMaxY = b[3];
    // src/Frame.cc (constructors): mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX), static_cast<float> on both
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}

// Frame with ncam cameras (1: the Nleft == -1 layout, keypoints in mvKeysUn; 2 / 4: the multi-camera layout)
void fill_frame(Frame &F, int ncam, const int *cam_off, const ref_kp *kps, const uint8_t *desc, const float *sf, int nsf) {
    F.N = cam_off[ncam];
    if (ncam == 1) {
        F.Nleft = F.Nright = F.Nsideleft = F.Nsideright = -1;
        for (int i = 0; i < F.N; i++) F.mvKeysUn.push_back(to_cv(kps[i]));
        F.mvKeys = F.mvKeysUn;
    } else {
        F.Nleft = cam_off[1] - cam_off[0]; F.Nright = cam_off[2] - cam_off[1];
        F.Nsideleft = ncam >= 4 ? cam_off[3] - cam_off[2] : -1; F.Nsideright = ncam >= 4 ? cam_off[4] - cam_off[3] : -1;
        for (int i = cam_off[0]; i < cam_off[1]; i++) F.mvKeys.push_back(to_cv(kps[i]));
        for (int i = cam_off[1]; i < cam_off[2]; i++) F.mvKeysRight.push_back(to_cv(kps[i]));
        if (ncam >= 4) {
            for (int i = cam_off[2]; i < cam_off[3]; i++) F.mvKeysSideLeft.push_back(to_cv(kps[i]));
            for (int i = cam_off[3]; i < cam_off[4]; i++) F.mvKeysSideRight.push_back(to_cv(kps[i]));
        }
        F.mvKeysUn = F.mvKeys;
    }
    F.mDescriptors = desc_mat(desc, F.N);
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f);
    F.mvDepth.assign(F.N, -1.f);
    F.mvScaleFactors.assign(sf, sf + nsf);
    F.AssignFeaturesToGrid();
}

#ifndef ORBX_GLUE_BUILD
class MatcherProbe : public ORBmatcher {
  public:
    using ORBmatcher::ORBmatcher;
    void three_maxima(std::vector<int> *h, int L, int &a, int &b, int &c) { ComputeThreeMaxima(h, L, a, b, c); }
};
#endif

}  // namespace

extern "C" {

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2577-2591)
int refm_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    cv::Mat ma(1, 32, CV_8U, (void *)a), mb(1, 32, CV_8U, (void *)b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

#ifndef ORBX_GLUE_BUILD
// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:2537-2573) on a histogram given by its bin sizes
void refm_three_maxima(const int *sizes, int L, int *ind3) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; i++) h[i].assign(sizes[i], 0);
    MatcherProbe m(0.6f, true);
    ind3[0] = ind3[1] = ind3[2] = -1;
    m.three_maxima(h.data(), L, ind3[0], ind3[1], ind3[2]);
}
#endif

// Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea (src/Frame.cc:541-582, 890-978) of one camera's keypoints
int refm_grid_query(const ref_kp *kps, int n, const float *bounds, float x, float y, float r, int minLevel, int maxLevel, int *out, int cap) {
    set_bounds(bounds);
    Frame F;
    const int cam_off[2] = {0, n};
    const float sf1[1] = {1.f};
    fill_frame(F, 1, cam_off, kps, std::vector<uint8_t>((size_t)n * 32 + 1).data(), sf1, 1);
    std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &, th, bFarPoints, thFarPoints) (src/ORBmatcher.cc:23-340); same
// arrays as orc_search_by_projection, `bounds` = {mnMinX, mnMaxX, mnMinY, mnMaxY} instead of prebuilt grids
int refm_search_by_projection(int ncam, const int *cam_off, const ref_kp *kps, const uint8_t *desc, const float *bounds, const float *sf, int nsf,
                              const uint8_t *occupied, const int *l2r, const int *r2l, const float *u_right, int nmp, const uint8_t *in_view,
                              const float *proj_x, const float *proj_y, const int *level, const float *view_cos, const float *proj_xr,
                              const float *depth, const uint8_t *bad, const uint8_t *has_obs, const uint8_t *mp_desc, float th, int bFarPoints,
                              float thFarPoints, float nnratio, int *assigned) {
    set_bounds(bounds);
    Frame F;
    fill_frame(F, ncam, cam_off, kps, desc, sf, nsf);
    const int N = F.N;
    MapPoint holder;                     // stands for "a map point with Observations() > 0 already sits here"
    holder.n_obs = 1;
    for (int i = 0; i < N; i++) if (occupied[i]) F.mvpMapPoints[i] = &holder;
    if (ncam >= 2) {
        F.mvLeftToRightMatch.assign(F.Nleft, -1); F.mvRightToLeftMatch.assign(F.Nright, -1);
        if (l2r) F.mvLeftToRightMatch.assign(l2r, l2r + F.Nleft);
        if (r2l) F.mvRightToLeftMatch.assign(r2l, r2l + F.Nright);
    }
    if (u_right && ncam == 1) F.mvuRight.assign(u_right, u_right + N);
    std::vector<MapPoint> mps(nmp);
    std::vector<MapPoint *> vp(nmp);
    for (int m = 0; m < nmp; m++) {
        MapPoint &P = mps[m];
        P.mbTrackInView = in_view[4 * m] != 0; P.mbTrackInViewR = in_view[4 * m + 1] != 0;
        P.mbTrackInViewSL = in_view[4 * m + 2] != 0; P.mbTrackInViewSR = in_view[4 * m + 3] != 0;
        P.mTrackProjX = proj_x[4 * m]; P.mTrackProjXR = ncam == 1 ? (proj_xr ? proj_xr[m] : 0.f) : proj_x[4 * m + 1];
        P.mTrackProjXSL = proj_x[4 * m + 2]; P.mTrackProjXSR = proj_x[4 * m + 3];
        P.mTrackProjY = proj_y[4 * m]; P.mTrackProjYR = proj_y[4 * m + 1]; P.mTrackProjYSL = proj_y[4 * m + 2]; P.mTrackProjYSR = proj_y[4 * m + 3];
        P.mnTrackScaleLevel = level[4 * m]; P.mnTrackScaleLevelR = level[4 * m + 1];
        P.mnTrackScaleLevelSL = level[4 * m + 2]; P.mnTrackScaleLevelSR = level[4 * m + 3];
        P.mTrackViewCos = view_cos[4 * m]; P.mTrackViewCosR = view_cos[4 * m + 1];
        P.mTrackViewCosSL = view_cos[4 * m + 2]; P.mTrackViewCosSR = view_cos[4 * m + 3];
        P.mTrackDepth = depth ? depth[m] : 0.f;
        P.bad = bad && bad[m];
        P.n_obs = has_obs ? has_obs[m] : 1;
        P.descriptor = desc_mat(mp_desc + (size_t)m * 32, 1);
        vp[m] = &P;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, vp, th, bFarPoints != 0, thFarPoints);
    for (int i = 0; i < N; i++) {
        MapPoint *p = F.mvpMapPoints[i];
        assigned[i] = (p && p != &holder) ? (int)(p - mps.data()) : -1;
    }
    return n;
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:895-1004); same arrays as orc_search_for_initialization
int refm_search_for_initialization(const ref_kp *kps1, const uint8_t *desc1, int n1, const ref_kp *kps2, const uint8_t *desc2, int n2,
                                   const float *bounds, float *vbPrevMatched, int windowSize, float nnratio, int checkOrientation, int *vnMatches12) {
    set_bounds(bounds);
    Frame F1, F2;
    const int o1[2] = {0, n1}, o2[2] = {0, n2};
    const float sf1[1] = {1.f};
    fill_frame(F1, 1, o1, kps1, desc1, sf1, 1);
    fill_frame(F2, 1, o2, kps2, desc2, sf1, 1);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(vbPrevMatched[2 * i], vbPrevMatched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, checkOrientation != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
    for (int i = 0; i < n1; i++) { vnMatches12[i] = m12[i]; vbPrevMatched[2 * i] = prev[i].x; vbPrevMatched[2 * i + 1] = prev[i].y; }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches) (src/ORBmatcher.cc:349-666); same arrays as orc_search_by_bow_frame
// plus the camera blocks of the KeyFrame's features (cam_off_kf, only the keypoint angles depend on them).  match_f[nF] = KeyFrame
// feature whose MapPoint ended up in vpMapPointMatches[i], or -1.
int refm_search_by_bow_frame(int nnKF, const int *nodesKF, const int *ptrKF, const int *idxKF, int nnF, const int *nodesF, const int *ptrF,
                             const int *idxF, const uint8_t *descKF, int nKF, const uint8_t *descF, int nF, const uint8_t *skipKF,
                             const float *angleKF, const float *angleF, int ncam, const int *camOffKF, const int *camOffF, float nnratio,
                             int checkOrientation, int *matchF) {
    const float bounds[4] = {0.f, 1000.f, 0.f, 1000.f};
    set_bounds(bounds);
    std::vector<ref_kp> kf(nKF), ff(nF);
    memset(kf.data(), 0, kf.size() * sizeof(ref_kp)); memset(ff.data(), 0, ff.size() * sizeof(ref_kp));
    for (int i = 0; i < nKF; i++) { kf[i].x = kf[i].y = 10.f; kf[i].angle = angleKF ? angleKF[i] : 0.f; }
    for (int i = 0; i < nF; i++) { ff[i].x = ff[i].y = 10.f; ff[i].angle = angleF ? angleF[i] : 0.f; }
    const float sf1[1] = {1.f};
    Frame F;
    fill_frame(F, ncam, camOffF, ff.data(), descF, sf1, 1);
    F.mvKeys = ncam == 1 ? F.mvKeysUn : F.mvKeys;
    for (int a = 0; a < nnF; a++)
        for (int j = ptrF[a]; j < ptrF[a + 1]; j++) F.mFeatVec[(unsigned)nodesF[a]].push_back((unsigned)idxF[j]);
    KeyFrame K;
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    K.N = nKF;
    K.mpCamera = &cams[0];
    if (ncam == 1) {
        for (int i = 0; i < nKF; i++) K.mvKeysUn.push_back(to_cv(kf[i]));
        K.mvKeys = K.mvKeysUn;
    } else {
        K.mpCamera2 = &cams[1];
        K.NLeft = camOffKF[1] - camOffKF[0]; K.NRight = camOffKF[2] - camOffKF[1];
        K.NSideLeft = ncam >= 4 ? camOffKF[3] - camOffKF[2] : -1; K.NSideRight = ncam >= 4 ? camOffKF[4] - camOffKF[3] : -1;
        for (int i = camOffKF[0]; i < camOffKF[1]; i++) K.mvKeys.push_back(to_cv(kf[i]));
        for (int i = camOffKF[1]; i < camOffKF[2]; i++) K.mvKeysRight.push_back(to_cv(kf[i]));
        if (ncam >= 4) {
            K.mpCamera3 = &cams[2]; K.mpCamera4 = &cams[3];
            for (int i = camOffKF[2]; i < camOffKF[3]; i++) K.mvKeysSideLeft.push_back(to_cv(kf[i]));
            for (int i = camOffKF[3]; i < camOffKF[4]; i++) K.mvKeysSideRight.push_back(to_cv(kf[i]));
        }
    }
    K.mDescriptors = desc_mat(descKF, nKF);
    std::vector<MapPoint> mps(nKF);
    K.map_points.assign(nKF, nullptr);
    for (int i = 0; i < nKF; i++) if (!skipKF[i]) K.map_points[i] = &mps[i];
    for (int a = 0; a < nnKF; a++)
        for (int j = ptrKF[a]; j < ptrKF[a + 1]; j++) K.mFeatVec[(unsigned)nodesKF[a]].push_back((unsigned)idxKF[j]);
    std::vector<MapPoint *> out;
    ORBmatcher matcher(nnratio, checkOrientation != 0);
    const int n = matcher.SearchByBoW(&K, F, out);
    for (int i = 0; i < nF; i++) matchF[i] = out[i] ? (int)(out[i] - mps.data()) : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:1006-1129); same arrays as
// orc_search_by_bow_keyframes.  match12[n1] = idx2 whose MapPoint ended up in vpMatches12[idx1], or -1.
int refm_search_by_bow_keyframes(int nn1, const int *nodes1, const int *ptr1, const int *idx1v, int nn2, const int *nodes2, const int *ptr2,
                                 const int *idx2v, const uint8_t *desc1, int n1, const uint8_t *desc2, int n2, const uint8_t *skip1,
                                 const uint8_t *skip2, const float *angle1, const float *angle2, float nnratio, int checkOrientation, int *match12) {
    KeyFrame K1, K2;
    std::vector<MapPoint> mp1(n1), mp2(n2);
    auto fill = [&](KeyFrame &K, int n, const uint8_t *desc, const uint8_t *skip, const float *angle, std::vector<MapPoint> &mps, int nn, const int *nodes,
                    const int *ptr, const int *idxv) {
        K.N = n;
        for (int i = 0; i < n; i++) K.mvKeysUn.push_back(cv::KeyPoint(10.f, 10.f, 31.f, angle ? angle[i] : 0.f, 0.f, 0, -1));
        K.mvKeys = K.mvKeysUn;
        K.mDescriptors = desc_mat(desc, n);
        K.map_points.assign(n, nullptr);
        for (int i = 0; i < n; i++) if (!skip || !skip[i]) K.map_points[i] = &mps[i];
        for (int a = 0; a < nn; a++)
            for (int j = ptr[a]; j < ptr[a + 1]; j++) K.mFeatVec[(unsigned)nodes[a]].push_back((unsigned)idxv[j]);
    };
    fill(K1, n1, desc1, skip1, angle1, mp1, nn1, nodes1, ptr1, idx1v);
    fill(K2, n2, desc2, skip2, angle2, mp2, nn2, nodes2, ptr2, idx2v);
    std::vector<MapPoint *> out;
    ORBmatcher matcher(nnratio, checkOrientation != 0);
    const int n = matcher.SearchByBoW(&K1, &K2, out);
    for (int i = 0; i < n1; i++) match12[i] = out[i] ? (int)(out[i] - mp2.data()) : -1;
    return n;
}

// ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th, cameraID = 0) (src/ORBmatcher.cc:1458-1640) on a monocular /
// rectified-stereo keyframe (NLeft == -1).  The keyframe sits at the origin with the pixel camera; map point q sits at (qx, qy, 1), so the
// reference projects it to (qx, qy), uses invz = 1 (ur = qx - bf), predicts the level the driver is given, and searches
// KeyFrame::GetFeaturesInArea (lifted from src/KeyFrame.cc).  EVERY keyframe feature already holds a map point with more observations than
// any query point: a query with bestDist <= TH_LOW then takes the branch `pMP->Replace(pMPinKF)` (:1630-1631), which leaves the keyframe
// untouched and records the chosen feature — best_idx[q] = that feature, or -1 when the point was not fused.  Returns nFused.
int refm_fuse(const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, const float *inv_sigma2, int nlev, const float *u_right,
              float bf, int nq, const uint8_t *valid, const float *qx, const float *qy, const int *level, const uint8_t *q_desc, float th, int *best_idx) {
    set_bounds(bounds);
    Frame F;                                                   // builds the grid the KeyFrame constructor copies (src/KeyFrame.cc:113-131)
    const int cam_off[2] = {0, n};
    fill_frame(F, 1, cam_off, kps, desc, sf, nlev);
    KeyFrame K;
    PixelCamera cam(0);
    K.N = n; K.mpCamera = &cam;
    K.mvKeysUn = F.mvKeysUn; K.mvKeys = F.mvKeysUn;
    K.mDescriptors = F.mDescriptors;
    K.mvScaleFactors.assign(sf, sf + nlev); K.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlev);
    K.mvuRight.assign(n, -1.f);
    if (u_right) K.mvuRight.assign(u_right, u_right + n);
    K.mbf = bf;
    K.mnMinX = (int)bounds[0]; K.mnMaxX = (int)bounds[1]; K.mnMinY = (int)bounds[2]; K.mnMaxY = (int)bounds[3];
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) K.mGrid[i][j] = F.mGrid[i][j];
    std::vector<MapPoint> holders(n), mps(nq);
    K.map_points.assign(n, nullptr);
    for (int i = 0; i < n; i++) { holders[i].n_obs = 1000; K.map_points[i] = &holders[i]; }
    std::vector<MapPoint *> vp(nq, nullptr);
    for (int q = 0; q < nq; q++) {
        if (!valid[q]) continue;
        MapPoint &P = mps[q];
        P.world_pos = Eigen::Vector3f(qx[q], qy[q], 1.f);
        const float d = P.world_pos.norm();
        P.normal = P.world_pos / d;                            // PO.dot(Pn) = dist3D >= 0.5 dist3D (:1557-1563)
        P.predicted_level = level[q];
        P.n_obs = 1;
        P.descriptor = desc_mat(q_desc + (size_t)q * 32, 1);
        vp[q] = &P;
    }
    ORBmatcher matcher(0.6f, true);
    const int nf = matcher.Fuse(&K, vp, th, 0);
    for (int q = 0; q < nq; q++) best_idx[q] = (valid[q] && mps[q].replaced_by) ? (int)(mps[q].replaced_by - holders.data()) : -1;
    return nf;
}

// ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th, 0) with every branch of its graph update (:1628-1645): keyframe as in refm_fuse, but
// kf_state[idx] chooses what the keyframe feature holds: 0 = a point with 1000 observations (the query is Replace()d by it), 1 = no point
// (AddObservation / AddMapPoint), 2 = a point with no observations (it is Replace()d by the query), 3 = a bad point (nothing happens, still
// counted).  dup_of[q] >= 0: query q is the SAME MapPoint object as query dup_of[q] (a point observed by two cameras appears twice in
// GetMapPointMatches()) — once fused it is in the keyframe (or bad) and the reference skips it at its second turn.
// Outputs: q_repl[q] = feature whose point replaced query q (-1 none; -2 - j: replaced by query j, which an earlier turn had added), kf_mp[idx] = query stored in the keyframe by AddMapPoint (-1),
// holder_repl[idx] = query that replaced the point the feature held (-1).  Returns nFused.
int refm_fuse_graph(const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, const float *inv_sigma2, int nlev,
                    const float *u_right, float bf, const uint8_t *kf_state, int nq, const uint8_t *valid, const int *dup_of, const float *qx, const float *qy,
                    const int *level, const uint8_t *q_desc, float th, int *q_repl, int *kf_mp, int *holder_repl) {
    set_bounds(bounds);
    Frame F;
    const int cam_off[2] = {0, n};
    fill_frame(F, 1, cam_off, kps, desc, sf, nlev);
    KeyFrame K;
    PixelCamera cam(0);
    K.N = n; K.mpCamera = &cam;
    K.mvKeysUn = F.mvKeysUn; K.mvKeys = F.mvKeysUn;
    K.mDescriptors = F.mDescriptors;
    K.mvScaleFactors.assign(sf, sf + nlev); K.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlev);
    K.mvuRight.assign(n, -1.f);
    if (u_right) K.mvuRight.assign(u_right, u_right + n);
    K.mbf = bf;
    K.mnMinX = (int)bounds[0]; K.mnMaxX = (int)bounds[1]; K.mnMinY = (int)bounds[2]; K.mnMaxY = (int)bounds[3];
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) K.mGrid[i][j] = F.mGrid[i][j];
    std::vector<MapPoint> holders(n), mps(nq);
    K.map_points.assign(n, nullptr);
    for (int i = 0; i < n; i++) {
        const int st = kf_state ? kf_state[i] : 0;
        holders[i].n_obs = st == 2 ? 0 : 1000; holders[i].bad = st == 3;
        if (st != 1) K.map_points[i] = &holders[i];
    }
    std::vector<MapPoint *> vp(nq, nullptr);
    for (int q = 0; q < nq; q++) {
        if (!valid[q]) continue;
        if (dup_of && dup_of[q] >= 0 && dup_of[q] < q && valid[dup_of[q]]) { vp[q] = vp[dup_of[q]]; continue; }
        MapPoint &P = mps[q];
        P.world_pos = Eigen::Vector3f(qx[q], qy[q], 1.f);
        P.normal = P.world_pos / P.world_pos.norm();
        P.predicted_level = level[q];
        P.n_obs = 1;
        P.descriptor = desc_mat(q_desc + (size_t)q * 32, 1);
        vp[q] = &P;
    }
    ORBmatcher matcher(0.6f, true);
    const int nf = matcher.Fuse(&K, vp, th, 0);
    for (int q = 0; q < nq; q++) {
        MapPoint *r = vp[q] == &mps[q] ? mps[q].replaced_by : nullptr;
        q_repl[q] = !r ? -1 : (r >= holders.data() && r < holders.data() + n) ? (int)(r - holders.data()) : -2 - (int)(r - mps.data());   // (-2 - j: by query j, added earlier)
    }
    for (int i = 0; i < n; i++) {
        MapPoint *m = K.map_points[i];
        kf_mp[i] = (m && m >= mps.data() && m < mps.data() + nq) ? (int)(m - mps.data()) : -1;
        MapPoint *r = holders[i].replaced_by;
        holder_repl[i] = (r && r >= mps.data() && r < mps.data() + nq) ? (int)(r - mps.data()) : -1;
    }
    return nf;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (src/ORBmatcher.cc:1985-2413), the tracking
// matcher.  Current frame as in refm_search_by_projection.  Last frame: npts keypoints in the Nleft == -1 layout; has_mp[i] = the point
// exists and is no outlier; its map point sits at (px[i], py[i], 1) in world = camera coordinates (Tcw = I, pixel camera), so the
// reference projects it to exactly (px[i], py[i]) in camera 0, to (px + trl[0], py + trl[1]) in the right camera (Trl = [I | trl]) and
// searches the side cameras around the LEFT projection (the reference's D6 behaviour, :2250-2271).  invzc = 1, so the rectified-stereo
// test uses ur = px - mbf.  gate: 0 = neither bForward nor bBackward, 1 = bForward, 2 = bBackward (set through the z of the last pose).
// assigned[idx] = i written to CurrentFrame.mvpMapPoints[idx] (-1 = untouched or cleared by the rotation filter; -3 = the keypoint still
// holds the unobserved point it held before the call, occupied[idx] == 2).
int refm_search_by_projection_last(int ncam, const int *cam_off, const ref_kp *kps, const uint8_t *desc, const float *bounds, const float *sf, int nsf,
                                   const uint8_t *occupied, const float *u_right, int npts, const uint8_t *has_mp, const float *px, const float *py,
                                   const float *trl, float mbf, const int *last_octave, const float *last_angle, const uint8_t *has_obs,
                                   const uint8_t *pt_desc, float th, int gate, int checkOrientation, int *assigned) {
    set_bounds(bounds);
    Frame Cur, Last;
    fill_frame(Cur, ncam, cam_off, kps, desc, sf, nsf);
    const int N = Cur.N;
    MapPoint holder, holder0;                                  // occupied[i] == 1: a point with observations; == 2: a point WITHOUT (claimable, :2072-2074)
    holder.n_obs = 1; holder0.n_obs = 0;
    for (int i = 0; i < N; i++) if (occupied && occupied[i]) Cur.mvpMapPoints[i] = occupied[i] == 2 ? &holder0 : &holder;
    if (u_right && ncam == 1) Cur.mvuRight.assign(u_right, u_right + N);
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    Cur.mpCamera = &cams[0]; Cur.mpCamera2 = ncam >= 2 ? &cams[1] : nullptr; Cur.mpCamera3 = ncam >= 4 ? &cams[2] : nullptr; Cur.mpCamera4 = ncam >= 4 ? &cams[3] : nullptr;
    Cur.mb = 0.5f; Cur.mbf = mbf;
    Cur.Trl = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(trl[0], trl[1], 0.f));
    // tlc = Tlw * twc = the last pose's translation (twc = 0): z = +1 -> bForward, -1 -> bBackward, 0 -> neither (mb = 0.5)
    Last.Tcw = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(0.f, 0.f, gate == 1 ? 1.f : gate == 2 ? -1.f : 0.f));
    Last.N = npts;
    std::vector<MapPoint> mps(npts);
    Last.mvpMapPoints.assign(npts, nullptr);
    Last.mvbOutlier.assign(npts, false);
    for (int i = 0; i < npts; i++) {
        Last.mvKeysUn.push_back(cv::KeyPoint(0.f, 0.f, 31.f, last_angle[i], 0.f, last_octave[i], -1));
        MapPoint &P = mps[i];
        P.world_pos = Eigen::Vector3f(px[i], py[i], 1.f);
        P.descriptor = desc_mat(pt_desc + (size_t)i * 32, 1);
        P.n_obs = has_obs ? has_obs[i] : 1;
        if (has_mp[i]) Last.mvpMapPoints[i] = &P;
    }
    Last.mvKeys = Last.mvKeysUn;
    ORBmatcher matcher(0.9f, checkOrientation != 0);
    const int n = matcher.SearchByProjection(Cur, Last, th, false);
    for (int i = 0; i < N; i++) {
        MapPoint *p = Cur.mvpMapPoints[i];
        assigned[i] = p == &holder0 ? -3 : (p && p != &holder) ? (int)(p - mps.data()) : -1;   // -3: still the unobserved point it held (not claimed, not cleared)
    }
    return n;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:1131-1456).
// FeatureVectors as CSR (node ids ascending); ncam = 1 (mono / rectified stereo: NLeft == -1), 2 or 4; has_mp[idx] = GetMapPoint(idx)
// != NULL; u_right (ncam == 1 only, may be NULL = all -1); ep = the epipole the reference computes (pKF2->mpCamera->project(T2w * Cw)):
// the driver places KF1's camera centre so that the pixel camera projects it to `ep`; sf2 = pKF2->mvScaleFactors.
// match12[n1] = idx2 or -1 (from vMatchedPairs).  Returns nmatches.
int refm_search_for_triangulation(int ncam, const int *cam_off1, const ref_kp *kps1, const uint8_t *desc1, const uint8_t *has_mp1, const float *u_right1,
                                  int nn1, const int *nodes1, const int *ptr1, const int *idx1v, const int *cam_off2, const ref_kp *kps2,
                                  const uint8_t *desc2, const uint8_t *has_mp2, const float *u_right2, int nn2, const int *nodes2, const int *ptr2,
                                  const int *idx2v, const float *ep, const float *sf2, int nsf, int bOnlyStereo, int bCoarse, int checkOrientation,
                                  int *match12) {
    KeyFrame K1, K2;
    PixelCamera c1[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    PixelCamera c2[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    MapPoint holder;
    auto fill = [&](KeyFrame &K, const int *cam_off, const ref_kp *kps, const uint8_t *desc, const uint8_t *has_mp, const float *u_right, int nn,
                    const int *nodes, const int *ptr, const int *idxv, PixelCamera *cams) {
        K.N = cam_off[ncam];
        if (ncam == 1) {
            for (int i = 0; i < K.N; i++) K.mvKeysUn.push_back(to_cv(kps[i]));
            K.mvKeys = K.mvKeysUn;
        } else {
            K.NLeft = cam_off[1] - cam_off[0]; K.NRight = cam_off[2] - cam_off[1];
            K.NSideLeft = ncam >= 4 ? cam_off[3] - cam_off[2] : -1; K.NSideRight = ncam >= 4 ? cam_off[4] - cam_off[3] : -1;
            for (int i = cam_off[0]; i < cam_off[1]; i++) K.mvKeys.push_back(to_cv(kps[i]));
            for (int i = cam_off[1]; i < cam_off[2]; i++) K.mvKeysRight.push_back(to_cv(kps[i]));
            if (ncam >= 4) {
                for (int i = cam_off[2]; i < cam_off[3]; i++) K.mvKeysSideLeft.push_back(to_cv(kps[i]));
                for (int i = cam_off[3]; i < cam_off[4]; i++) K.mvKeysSideRight.push_back(to_cv(kps[i]));
            }
        }
        K.mDescriptors = desc_mat(desc, K.N);
        K.map_points.assign(K.N, nullptr);
        for (int i = 0; i < K.N; i++) if (has_mp[i]) K.map_points[i] = &holder;
        K.mvuRight.assign(K.N, -1.f);
        if (u_right && ncam == 1) K.mvuRight.assign(u_right, u_right + K.N);
        K.mvScaleFactors.assign(sf2, sf2 + nsf);
        K.mvLevelSigma2.assign(nsf, 1.f);
        for (int i = 0; i < nsf; i++) K.mvLevelSigma2[i] = sf2[i] * sf2[i];
        K.mpCamera = &cams[0];
        if (ncam >= 2) K.mpCamera2 = &cams[1];
        if (ncam >= 4) { K.mpCamera3 = &cams[2]; K.mpCamera4 = &cams[3]; }
        for (int a = 0; a < nn; a++)
            for (int j = ptr[a]; j < ptr[a + 1]; j++) K.mFeatVec[(unsigned)nodes[a]].push_back((unsigned)idxv[j]);
    };
    fill(K1, cam_off1, kps1, desc1, has_mp1, u_right1, nn1, nodes1, ptr1, idx1v, c1);
    fill(K2, cam_off2, kps2, desc2, has_mp2, u_right2, nn2, nodes2, ptr2, idx2v, c2);
    // KF2 at the origin; KF1's camera centre Cw = (ep.x, ep.y, 1): T1w = [I | -Cw], so C2 = T2w * Cw = Cw and project(C2) = ep exactly
    K1.Tcw = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(-ep[0], -ep[1], -1.f));
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(0.6f, checkOrientation != 0);
    const int n = matcher.SearchForTriangulation(&K1, &K2, pairs, bOnlyStereo != 0, bCoarse != 0);
    for (int i = 0; i < K1.N; i++) match12[i] = -1;
    for (auto &p : pairs) match12[p.first] = (int)p.second;
    return n;
}

// ORBmatcher::SearchForTriangulation with REAL pinhole geometry: cameras cam[c] = {fx, fy, cx, cy} (the same rig in both keyframes),
// identity rotations, KF2 at the origin, KF1's pose translation t1w, right / side-left / side-right cameras at the translations rel[0..2]
// from the left one.  The reference computes the epipole, the eleven (R12, t12) pairs and — through Pinhole::epipolarConstrain, its own
// code — the fundamental matrices itself.  sigma2[nsf] = mvLevelSigma2 of both keyframes.  Everything else as refm_search_for_triangulation.
int refm_search_for_triangulation_pinhole(int ncam, const int *cam_off1, const ref_kp *kps1, const uint8_t *desc1, const uint8_t *has_mp1,
                                          const float *u_right1, int nn1, const int *nodes1, const int *ptr1, const int *idx1v, const int *cam_off2,
                                          const ref_kp *kps2, const uint8_t *desc2, const uint8_t *has_mp2, const float *u_right2, int nn2,
                                          const int *nodes2, const int *ptr2, const int *idx2v, const float *cam, const float *t1w, const float *rel,
                                          const float *sf2, const float *sigma2, int nsf, int bOnlyStereo, int bCoarse, int checkOrientation,
                                          int *match12) {
    KeyFrame K1, K2;
    std::vector<Pinhole> c1, c2;
    for (int c = 0; c < 4; c++) { c1.push_back(Pinhole(cam[4 * c], cam[4 * c + 1], cam[4 * c + 2], cam[4 * c + 3])); c2.push_back(c1.back()); }
    MapPoint holder;
    auto fill = [&](KeyFrame &K, const int *cam_off, const ref_kp *kps, const uint8_t *desc, const uint8_t *has_mp, const float *u_right, int nn,
                    const int *nodes, const int *ptr, const int *idxv, std::vector<Pinhole> &cams) {
        K.N = cam_off[ncam];
        if (ncam == 1) {
            for (int i = 0; i < K.N; i++) K.mvKeysUn.push_back(to_cv(kps[i]));
            K.mvKeys = K.mvKeysUn;
        } else {
            K.NLeft = cam_off[1] - cam_off[0]; K.NRight = cam_off[2] - cam_off[1];
            K.NSideLeft = ncam >= 4 ? cam_off[3] - cam_off[2] : -1; K.NSideRight = ncam >= 4 ? cam_off[4] - cam_off[3] : -1;
            for (int i = cam_off[0]; i < cam_off[1]; i++) K.mvKeys.push_back(to_cv(kps[i]));
            for (int i = cam_off[1]; i < cam_off[2]; i++) K.mvKeysRight.push_back(to_cv(kps[i]));
            if (ncam >= 4) {
                for (int i = cam_off[2]; i < cam_off[3]; i++) K.mvKeysSideLeft.push_back(to_cv(kps[i]));
                for (int i = cam_off[3]; i < cam_off[4]; i++) K.mvKeysSideRight.push_back(to_cv(kps[i]));
            }
        }
        K.mDescriptors = desc_mat(desc, K.N);
        K.map_points.assign(K.N, nullptr);
        for (int i = 0; i < K.N; i++) if (has_mp[i]) K.map_points[i] = &holder;
        K.mvuRight.assign(K.N, -1.f);
        if (u_right && ncam == 1) K.mvuRight.assign(u_right, u_right + K.N);
        K.mvScaleFactors.assign(sf2, sf2 + nsf);
        K.mvLevelSigma2.assign(sigma2, sigma2 + nsf);
        K.mpCamera = &cams[0];
        if (ncam >= 2) K.mpCamera2 = &cams[1];
        if (ncam >= 4) { K.mpCamera3 = &cams[2]; K.mpCamera4 = &cams[3]; }
        K.Trl = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(rel[0], rel[1], rel[2]));
        K.Tsll = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(rel[3], rel[4], rel[5]));
        K.Tsrl = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(rel[6], rel[7], rel[8]));
        for (int a = 0; a < nn; a++)
            for (int j = ptr[a]; j < ptr[a + 1]; j++) K.mFeatVec[(unsigned)nodes[a]].push_back((unsigned)idxv[j]);
    };
    fill(K1, cam_off1, kps1, desc1, has_mp1, u_right1, nn1, nodes1, ptr1, idx1v, c1);
    fill(K2, cam_off2, kps2, desc2, has_mp2, u_right2, nn2, nodes2, ptr2, idx2v, c2);
    K1.Tcw = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(t1w[0], t1w[1], t1w[2]));
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(0.6f, checkOrientation != 0);
    const int n = matcher.SearchForTriangulation(&K1, &K2, pairs, bOnlyStereo != 0, bCoarse != 0);
    for (int i = 0; i < K1.N; i++) match12[i] = -1;
    for (auto &pr : pairs) match12[pr.first] = (int)pr.second;
    return n;
}

// the pixel camera's epipolar predicate, for the oracle-side callback of the same test
// KeyFrame with one camera block (NLeft == -1) over n keypoints: grid as the KeyFrame constructor copies it from the Frame
// (src/KeyFrame.cc:113-131); the other cameras' grids exist and are empty
static void fill_keyframe(KeyFrame &K, PixelCamera *cams, const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, int nlev) {
    Frame F;
    const int cam_off[2] = {0, n};
    fill_frame(F, 1, cam_off, kps, desc, sf, nlev);
    K.N = n;
    K.mpCamera = &cams[0]; K.mpCamera2 = &cams[1]; K.mpCamera3 = &cams[2]; K.mpCamera4 = &cams[3];
    K.mvKeysUn = F.mvKeysUn; K.mvKeys = F.mvKeysUn;
    K.mDescriptors = F.mDescriptors;
    K.mvScaleFactors.assign(sf, sf + nlev);
    K.mvuRight.assign(n, -1.f);
    K.mnMinX = (int)bounds[0]; K.mnMaxX = (int)bounds[1]; K.mnMinY = (int)bounds[2]; K.mnMaxY = (int)bounds[3];
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    K.mGridRight = K.mGrid; K.mGridSideLeft = K.mGrid; K.mGridSideRight = K.mGrid;
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) K.mGrid[i][j] = F.mGrid[i][j];
    K.map_points.assign(n, nullptr);
}

// a map point that projects to (x, y) with every geometric gate open: world = camera coordinates (identity poses), z = 1, normal = PO / |PO|
static void place_point(MapPoint &P, float x, float y, int level, const uint8_t *d) {
    P.world_pos = Eigen::Vector3f(x, y, 1.f);
    P.normal = P.world_pos / P.world_pos.norm();
    P.predicted_level = level;
    P.n_obs = 1;
    P.descriptor = desc_mat(d, 1);
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:2415-2535), the relocalisation matcher.  Current frame: one camera block, `occupied` keypoints hold a map point.
// KeyFrame: npts map points (has_mp), point i at (px, py, 1) with predicted level level[i], its keypoint angle kf_angle[i];
// already_found[i] = the point is in sAlreadyFound.  assigned[idx] = i written to CurrentFrame.mvpMapPoints[idx] (-1: untouched / cleared).
int refm_search_by_projection_reloc(const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, int nsf, const uint8_t *occupied,
                                    int npts, const uint8_t *has_mp, const uint8_t *already_found, const float *px, const float *py, const int *level,
                                    const float *kf_angle, const uint8_t *pt_desc, float th, int ORBdist, int checkOrientation, int *assigned) {
    set_bounds(bounds);
    Frame Cur;
    const int cam_off[2] = {0, n};
    fill_frame(Cur, 1, cam_off, kps, desc, sf, nsf);
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    Cur.mpCamera = &cams[0];
    MapPoint holder;
    holder.n_obs = 1;
    for (int i = 0; i < n; i++) if (occupied && occupied[i]) Cur.mvpMapPoints[i] = &holder;
    KeyFrame K;
    K.N = npts;
    std::vector<MapPoint> mps(npts);
    K.map_points.assign(npts, nullptr);
    std::set<MapPoint *> found;
    for (int i = 0; i < npts; i++) {
        K.mvKeysUn.push_back(cv::KeyPoint(0.f, 0.f, 31.f, kf_angle[i], 0.f, 0, -1));
        if (!has_mp[i]) continue;
        place_point(mps[i], px[i], py[i], level[i], pt_desc + (size_t)i * 32);
        K.map_points[i] = &mps[i];
        if (already_found && already_found[i]) found.insert(&mps[i]);
    }
    ORBmatcher matcher(0.9f, checkOrientation != 0);
    const int nm = matcher.SearchByProjection(Cur, &K, found, th, ORBdist);
    for (int i = 0; i < n; i++) {
        MapPoint *m = Cur.mvpMapPoints[i];
        assigned[i] = (m && m != &holder) ? (int)(m - mps.data()) : -1;
    }
    return nm;
}

// ORBmatcher::SearchByProjection(KeyFrame *pKF, Sim3f &Scw, vpPoints, vpMatched, th, ratioHamming, cameraID = 0) (src/ORBmatcher.cc:668-777)
// and, with with_kfs != 0, the overload that also records the points' keyframes (:779-893).  Scw = identity.  matched_in[idx] = vpMatched
// already holds a point there; candidate point i at (px, py, 1); skip[i] = the point is bad or already in vpMatched.
// assigned[idx] = i for the entries the call wrote (-1 elsewhere); matched_kf_ok = every written entry of vpMatchedKF is the point's keyframe.
int refm_search_by_projection_sim3(const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, int nsf, const uint8_t *matched_in,
                                   int npts, const uint8_t *skip, const float *px, const float *py, const int *level, const uint8_t *pt_desc, int th,
                                   float ratioHamming, int with_kfs, int *assigned, int *matched_kf_ok) {
    set_bounds(bounds);
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    KeyFrame K;
    fill_keyframe(K, cams, kps, desc, n, bounds, sf, nsf);
    std::vector<MapPoint> mps(npts), holders(n);
    std::vector<KeyFrame> kfs(npts);
    std::vector<MapPoint *> vpPoints(npts), vpMatched(n, nullptr);
    std::vector<KeyFrame *> vpPointsKFs(npts), vpMatchedKF(n, nullptr);
    for (int i = 0; i < n; i++) if (matched_in && matched_in[i]) vpMatched[i] = &holders[i];
    int nskip_found = 0;
    for (int i = 0; i < npts; i++) {
        place_point(mps[i], px[i], py[i], level[i], pt_desc + (size_t)i * 32);
        vpPoints[i] = &mps[i]; vpPointsKFs[i] = &kfs[i];
        if (skip && skip[i] == 1) mps[i].bad = true;
        if (skip && skip[i] == 2) { // already found: the point itself sits in vpMatched (at a keypoint that is occupied anyway)
            for (int k = nskip_found; k < n; k++) if (vpMatched[k]) { vpMatched[k] = &mps[i]; nskip_found = k + 1; break; }
        }
    }
    std::vector<MapPoint *> before = vpMatched;
    Sophus::Sim3f Scw;
    ORBmatcher matcher(0.9f, true);
    const int nm = with_kfs ? matcher.SearchByProjection(&K, Scw, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratioHamming, 0)
                            : matcher.SearchByProjection(&K, Scw, vpPoints, vpMatched, th, ratioHamming, 0);
    int ok = 1;
    for (int i = 0; i < n; i++) {
        assigned[i] = (vpMatched[i] && vpMatched[i] != before[i]) ? (int)(vpMatched[i] - mps.data()) : -1;
        if (with_kfs && assigned[i] >= 0 && vpMatchedKF[i] != &kfs[assigned[i]]) ok = 0;
    }
    if (matched_kf_ok) *matched_kf_ok = ok;
    return nm;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (src/ORBmatcher.cc:1771-1983): T1w = T2w = I, S12 = identity, fx = fy = 1,
// cx = cy = 0, so map point i1 of KF1 at (u1, v1, 1) projects to (u1, v1) in KF2 and vice versa.  pre12[i1] = keypoint of KF2 that
// vpMatches12[i1] already points to on entry (-1: none; -2: a point that is not in KF2; >= n2: an index the reference ignores, :1802).  match12[i1] = keypoint of KF2 the call matched (-1: none / was pre-matched).
int refm_search_by_sim3(const ref_kp *kps1, const uint8_t *desc1, int n1, const ref_kp *kps2, const uint8_t *desc2, int n2, const float *bounds,
                        const float *sf, int nsf, const uint8_t *has_mp1, const float *u1, const float *v1, const int *level1, const uint8_t *mpdesc1,
                        const uint8_t *has_mp2, const float *u2, const float *v2, const int *level2, const uint8_t *mpdesc2, const int *pre12, float th,
                        int *match12) {
    set_bounds(bounds);
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    KeyFrame K1, K2;
    fill_keyframe(K1, cams, kps1, desc1, n1, bounds, sf, nsf);
    fill_keyframe(K2, cams, kps2, desc2, n2, bounds, sf, nsf);
    std::vector<MapPoint> mp1(n1), mp2(n2), pre(n1);
    for (int i = 0; i < n1; i++) if (has_mp1[i]) { place_point(mp1[i], u1[i], v1[i], level1[i], mpdesc1 + (size_t)i * 32); K1.map_points[i] = &mp1[i]; }
    for (int i = 0; i < n2; i++) if (has_mp2[i]) { place_point(mp2[i], u2[i], v2[i], level2[i], mpdesc2 + (size_t)i * 32); K2.map_points[i] = &mp2[i]; }
    std::vector<MapPoint *> vpMatches12(n1, nullptr);
    for (int i = 0; i < n1; i++)
        if (pre12 && pre12[i] != -1) { if (pre12[i] >= 0) pre[i].index_in_kf[&K2] = pre12[i]; vpMatches12[i] = &pre[i]; }
    ORBmatcher matcher(0.75f, true);
    const Sophus::Sim3f S12;
    const int nf = matcher.SearchBySim3(&K1, &K2, vpMatches12, S12, th);
    for (int i = 0; i < n1; i++) {
        MapPoint *m = vpMatches12[i];
        match12[i] = (m && m >= mp2.data() && m < mp2.data() + n2) ? (int)(m - mp2.data()) : -1;
    }
    return nf;
}

// ORBmatcher::Fuse(KeyFrame *pKF, Sim3f &Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1648-1769), the loop-closing fuse: Scw =
// identity, the keyframe's keypoints with has_mp hold a map point (-> vpReplacePoint), the others get the point added.  Cameras 1-3 of the
// keyframe exist and are empty (the reference walks all four).  best_idx[q] = keypoint the search of point q ended on (-1: none).
int refm_fuse_sim3(const ref_kp *kps, const uint8_t *desc, int n, const float *bounds, const float *sf, int nlev, const uint8_t *has_mp, int nq,
                   const uint8_t *valid, const float *qx, const float *qy, const int *level, const uint8_t *q_desc, float th, int *best_idx) {
    set_bounds(bounds);
    PixelCamera cams[4] = {PixelCamera(0), PixelCamera(1), PixelCamera(2), PixelCamera(3)};
    KeyFrame K;
    fill_keyframe(K, cams, kps, desc, n, bounds, sf, nlev);
    std::vector<MapPoint> holders(n), mps(nq);
    for (int i = 0; i < n; i++) if (!has_mp || has_mp[i]) { holders[i].n_obs = 3; K.map_points[i] = &holders[i]; }
    std::vector<MapPoint *> vp(nq), rep(nq, nullptr);
    for (int q = 0; q < nq; q++) {
        place_point(mps[q], qx[q], qy[q], level[q], q_desc + (size_t)q * 32);
        mps[q].n_obs = 0;
        if (!valid[q]) mps[q].bad = true;
        vp[q] = &mps[q];
    }
    Sophus::Sim3f Scw;
    ORBmatcher matcher(0.8f, true);
    const int nf = matcher.Fuse(&K, Scw, vp, th, rep);
    for (int q = 0; q < nq; q++) {
        best_idx[q] = -1;
        MapPoint *r = rep[q];
        if (r && r >= holders.data() && r < holders.data() + n) best_idx[q] = (int)(r - holders.data());
        else if (r) best_idx[q] = r->index_in_kf[&K];                   // replaced by a point an earlier search added at that keypoint
        else if (mps[q].index_in_kf.count(&K)) best_idx[q] = mps[q].index_in_kf[&K];   // added
    }
    return nf;
}

int refm_pair_hash(float x1, float y1, float x2, float y2, int id1, int id2) { return PixelCamera::pair_hash(x1, y1, x2, y2, id1, id2) ? 1 : 0; }

}  // extern "C"
