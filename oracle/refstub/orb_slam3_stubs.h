// oracle/refstub/orb_slam3_stubs.h — TEST INFRASTRUCTURE.
// What the UNMODIFIED reference file /root/reference/src/ORBmatcher.cc needs from the rest of ORB-SLAM3 in order to compile here
// (oracle/Makefile: -DFRAME_H -DKEYFRAME_H -DMAPPOINT_H switches the bodies of the real include/{Frame,KeyFrame,MapPoint}.h off, this
// header is force-included instead): minimal Eigen / Sophus value types, an abstract camera, and Frame / KeyFrame / MapPoint holding
// exactly the data members and accessors ORBmatcher.cc touches, with the reference's names and meanings (include/Frame.h:110-340,
// include/KeyFrame.h, include/MapPoint.h).  The grid functions Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid and
// KeyFrame::GetFeaturesInArea are NOT restated: oracle/Makefile lifts their bodies from src/Frame.cc / src/KeyFrame.cc at build time
// into oracle/_ref/gen/ (never into the repo) and oracle/ref_matcher_driver.cpp includes them.
// The matchers' geometry (poses, projections, epipolar tests) is an injected dependency: tests use identity poses and a "pixel
// camera" whose project() returns (x, y) of the 3-D point, so that projections are exact and the reference's selection logic —
// candidate order, gates, best / second bookkeeping, claims, rotation histogram — is what gets compared.
#pragma once
#include <assert.h>
#include <math.h>

#include <list>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// include/ORBmatcher.h names `vector` / `pair` unqualified: the reference's own headers pull namespace std into the global scope
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36, reached through include/Frame.h -> include/ORBVocabulary.h)
using namespace std;

namespace Eigen {
template <int N> struct Vec {
    float v[N];
    Vec() { for (int i = 0; i < N; i++) v[i] = 0.f; }
    Vec(float a, float b) { static_assert(N == 2, ""); v[0] = a; v[1] = b; }
    Vec(float a, float b, float c) { static_assert(N == 3, ""); v[0] = a; v[1] = b; v[2] = c; }
    float &operator()(int i) { return v[i]; }
    const float &operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    const float &operator[](int i) const { return v[i]; }
    float dot(const Vec &o) const { float s = 0.f; for (int i = 0; i < N; i++) s += v[i] * o.v[i]; return s; }
    float squaredNorm() const { return dot(*this); }
    float norm() const { return sqrtf(dot(*this)); }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { static_assert(N >= 3, ""); return v[2]; }
};
template <int N> static inline Vec<N> operator+(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
template <int N> static inline Vec<N> operator-(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
template <int N> static inline Vec<N> operator-(const Vec<N> &a) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = -a.v[i]; return r; }
template <int N> static inline Vec<N> operator*(const Vec<N> &a, float s) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] * s; return r; }
template <int N> static inline Vec<N> operator*(float s, const Vec<N> &a) { return a * s; }
template <int N> static inline Vec<N> operator/(const Vec<N> &a, float s) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] / s; return r; }
typedef Vec<2> Vector2f;
typedef Vec<3> Vector3f;
struct Matrix3f {
    float m[3][3];
    Matrix3f() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = 0.f; }
    static Matrix3f Identity() { Matrix3f r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.f; return r; }
    float &operator()(int i, int j) { return m[i][j]; }
    const float &operator()(int i, int j) const { return m[i][j]; }
    Matrix3f transpose() const { Matrix3f r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i]; return r; }
    Matrix3f inverse() const {   // adjugate / determinant (what the harness cameras' fundamental matrices are built with, on both sides of a comparison)
        Matrix3f r;
        const float c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2], c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
        const float det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02, id = 1.f / det;
        r.m[0][0] = c00 * id; r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
        r.m[1][0] = c01 * id; r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
        r.m[2][0] = c02 * id; r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
        return r;
    }
};
static inline Vector3f operator*(const Matrix3f &A, const Vector3f &x) {
    Vector3f r;
    for (int i = 0; i < 3; i++) r.v[i] = A.m[i][0] * x.v[0] + A.m[i][1] * x.v[1] + A.m[i][2] * x.v[2];
    return r;
}
static inline Matrix3f operator*(const Matrix3f &A, const Matrix3f &B) {
    Matrix3f r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return r;
}
static inline Matrix3f operator*(const Matrix3f &A, float s) { Matrix3f r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][j] * s; return r; }
}  // namespace Eigen

namespace Sophus {
// rigid transform x -> R x + t and similarity x -> s R x + t, with the handful of members the matchers call
template <class T> class SE3 {
  public:
    SE3() : R_(Eigen::Matrix3f::Identity()) {}
    SE3(const Eigen::Matrix3f &R, const Eigen::Vector3f &t) : R_(R), t_(t) {}
    Eigen::Matrix3f rotationMatrix() const { return R_; }
    Eigen::Vector3f translation() const { return t_; }
    SE3 inverse() const { Eigen::Matrix3f Rt = R_.transpose(); return SE3(Rt, -(Rt * t_)); }
    Eigen::Vector3f operator*(const Eigen::Vector3f &x) const { return R_ * x + t_; }
    SE3 operator*(const SE3 &o) const { return SE3(R_ * o.R_, R_ * o.t_ + t_); }
  private:
    Eigen::Matrix3f R_;
    Eigen::Vector3f t_;
};
typedef SE3<float> SE3f;
template <class T> class SO3 {
  public:
    static Eigen::Matrix3f hat(const Eigen::Vector3f &w) {   // the cross-product matrix
        Eigen::Matrix3f r;
        r(0, 1) = -w(2); r(0, 2) = w(1); r(1, 0) = w(2); r(1, 2) = -w(0); r(2, 0) = -w(1); r(2, 1) = w(0);
        return r;
    }
};
typedef SO3<float> SO3f;
template <class T> class Sim3 {
  public:
    Sim3() : R_(Eigen::Matrix3f::Identity()), s_(1.f) {}
    Sim3(float s, const Eigen::Matrix3f &R, const Eigen::Vector3f &t) : R_(R), t_(t), s_(s) {}
    Eigen::Matrix3f rotationMatrix() const { return R_; }
    Eigen::Vector3f translation() const { return t_; }
    float scale() const { return s_; }
    Sim3 inverse() const { Eigen::Matrix3f Rt = R_.transpose(); return Sim3(1.f / s_, Rt, -((Rt * t_) * (1.f / s_))); }
    Eigen::Vector3f operator*(const Eigen::Vector3f &x) const { return (R_ * x) * s_ + t_; }
  private:
    Eigen::Matrix3f R_;
    Eigen::Vector3f t_;
    float s_;
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

class KeyFrame;
class Frame;

// include/CameraModels/GeometricCamera.h:25-103 — the three virtuals the matchers call
class GeometricCamera {
  public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f &v3D) = 0;
    virtual bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                                   const Eigen::Vector3f &t12, const float sigmaLevel, const float unc) = 0;
    virtual Eigen::Matrix3f toK_() { return Eigen::Matrix3f::Identity(); }   // (pure virtual in the reference)
    float getParameter(const int i) { return mvParameters[i]; }
    unsigned int GetType() { return mnType; }
    const static unsigned int CAM_PINHOLE = 0;
    const static unsigned int CAM_FISHEYE = 1;
  protected:
    std::vector<float> mvParameters;
    unsigned int mnType = CAM_PINHOLE;
};

// include/MapPoint.h — tracking fields (public data members in the reference too) and the accessors ORBmatcher.cc calls
class MapPoint {
  public:
    MapPoint() {}
    // SearchByProjection(Frame, MapPoints) inputs, written by Frame::isInFrustum in the reference (include/MapPoint.h:133-160)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    float mTrackProjXSL = 0, mTrackProjYSL = 0, mTrackProjXSR = 0, mTrackProjYSR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false, mbTrackInViewSL = false, mbTrackInViewSR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0, mnTrackScaleLevelSL = 0, mnTrackScaleLevelSR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0, mTrackViewCosSL = 0, mTrackViewCosSR = 0;
    long unsigned int mnLastFrameSeen = 0, mnFuseCandidateForKF = 0;
    long unsigned int mnId = 0;

    // test-side state
    cv::Mat descriptor;                       // 1 x 32 CV_8U
    Eigen::Vector3f world_pos, normal;
    bool bad = false;
    int n_obs = 0;
    int predicted_level = 0;
    float min_dist = 0.f, max_dist = 1e30f;
    std::map<KeyFrame *, int> index_in_kf;    // GetIndexInKeyFrame / IsInKeyFrame
    MapPoint *replaced_by = nullptr;

    cv::Mat GetDescriptor() { return descriptor.clone(); }                 // src/MapPoint.cc:485-488
    bool isBad() { return bad; }
    int Observations() { return n_obs; }
    Eigen::Vector3f GetWorldPos() { return world_pos; }
    Eigen::Vector3f GetNormal() { return normal; }
    float GetMinDistanceInvariance() { return min_dist; }
    float GetMaxDistanceInvariance() { return max_dist; }
    int PredictScale(const float &, KeyFrame *) { return predicted_level; }
    int PredictScale(const float &, Frame *) { return predicted_level; }
    std::tuple<int, int, int, int> GetIndexInKeyFrame(KeyFrame *kf) {   // src/MapPoint.cc:490-496 (left, right, side-left, side-right)
        auto it = index_in_kf.find(kf);
        return it == index_in_kf.end() ? std::tuple<int, int, int, int>(-1, -1, -1, -1) : std::tuple<int, int, int, int>(it->second, -1, -1, -1);
    }
    bool IsInKeyFrame(KeyFrame *kf) { return index_in_kf.count(kf) != 0; }
    void AddObservation(KeyFrame *kf, int idx) { index_in_kf[kf] = idx; n_obs++; }
    void Replace(MapPoint *other) { replaced_by = other; bad = true; }
};

// include/Frame.h — data members with the reference's names
class Frame {
  public:
    Frame() {}
    int N = 0;
    int Nleft = -1, Nright = -1, Nsideleft = -1, Nsideright = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysSideLeft, mvKeysSideRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mb = 0, mbf = 0;
    static float fx, fy, cx, cy, invfx, invfy;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr, *mpCamera3 = nullptr, *mpCamera4 = nullptr;
    DBoW2::FeatureVector mFeatVec;
    long unsigned int mnId = 0;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridSideLeft[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridSideRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    Sophus::SE3f Tcw, Trl, Tsll, Tsrl;

    // src/Frame.cc:541-582, 890-967, 969-978 — bodies lifted from the reference at build time
    void AssignFeaturesToGrid();
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1,
                                          const int selectedCamera = 0) const;
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);

    Sophus::SE3f GetPose() const { return Tcw; }
    Sophus::SE3f GetRelativePoseTrl() { return Trl; }
    Sophus::SE3f GetRelativePoseTsll() { return Tsll; }
    Sophus::SE3f GetRelativePoseTsrl() { return Tsrl; }
    bool isInFrustum(MapPoint *, float) { return true; }
};

// include/KeyFrame.h
class KeyFrame {
  public:
    KeyFrame() {}
    int N = 0;
    int NLeft = -1, NRight = -1, NSideLeft = -1, NSideRight = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysSideLeft, mvKeysSideRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr, *mpCamera3 = nullptr, *mpCamera4 = nullptr;
    DBoW2::FeatureVector mFeatVec;
    long unsigned int mnId = 0;
    std::vector<MapPoint *> map_points;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight, mGridSideLeft, mGridSideRight;
    Sophus::SE3f Tcw, Trl, Tsll, Tsrl;      // camera <- world of the left camera; right / side cameras <- left camera

    // src/KeyFrame.cc:771-834 — body lifted from the reference at build time
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int selectedCamera = 0) const;

    std::vector<MapPoint *> GetMapPointMatches() { return map_points; }
    MapPoint *GetMapPoint(const size_t &idx) { return map_points[idx]; }
    std::set<MapPoint *> GetMapPoints() { std::set<MapPoint *> s; for (auto *p : map_points) if (p && !p->isBad()) s.insert(p); return s; }
    void AddMapPoint(MapPoint *p, const size_t &idx) { map_points[idx] = p; }
    bool IsInImage(const float &x, const float &y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }   // include/KeyFrame.h
    bool isBad() { return false; }
    GeometricCamera *GetCamera(int id) { return id == 0 ? mpCamera : id == 1 ? mpCamera2 : id == 2 ? mpCamera3 : mpCamera4; }
    int kpIdxToCamIdx(const int kpIdx) const {                       // src/KeyFrame.cc:1329-1335
        return (NLeft == -1) ? 0 : (kpIdx < NLeft) ? 0 : (kpIdx < NLeft + NRight) ? 1 : (kpIdx < NLeft + NRight + NSideLeft) ? 2 : 3;
    }
    const cv::KeyPoint &GetKey(const int kpIdx) const {              // src/KeyFrame.cc:1312-1327
        switch (kpIdxToCamIdx(kpIdx)) {
            case 0: return mvKeys[kpIdx];
            case 1: return mvKeysRight[kpIdx - NLeft];
            case 2: return mvKeysSideLeft[kpIdx - NLeft - NRight];
            default: return mvKeysSideRight[kpIdx - NLeft - NRight - NSideLeft];
        }
    }
    Sophus::SE3f GetPose() { return Tcw; }
    Sophus::SE3f GetPoseInverse() { return Tcw.inverse(); }
    Eigen::Vector3f GetCameraCenter() { return Tcw.inverse().translation(); }
    Sophus::SE3f GetRightPose() { return Trl * Tcw; }
    Sophus::SE3f GetRightPoseInverse() { return (Trl * Tcw).inverse(); }
    Eigen::Vector3f GetRightCameraCenter() { return (Trl * Tcw).inverse().translation(); }
    Sophus::SE3f GetSideLeftPose() { return Tsll * Tcw; }
    Sophus::SE3f GetSideLeftPoseInverse() { return (Tsll * Tcw).inverse(); }
    Eigen::Vector3f GetSideLeftCameraCenter() { return (Tsll * Tcw).inverse().translation(); }
    Sophus::SE3f GetSideRightPose() { return Tsrl * Tcw; }
    Sophus::SE3f GetSideRightPoseInverse() { return (Tsrl * Tcw).inverse(); }
    Eigen::Vector3f GetSideRightCameraCenter() { return (Tsrl * Tcw).inverse().translation(); }
    Sophus::SE3f GetRelativePoseTrl() { return Trl; }
    Sophus::SE3f GetRelativePoseTsll() { return Tsll; }
    Sophus::SE3f GetRelativePoseTsrl() { return Tsrl; }
};

}  // namespace ORB_SLAM3
