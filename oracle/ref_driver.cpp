// oracle/ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the UNMODIFIED reference class
// ORB_SLAM3::ORBextractor (compiled by oracle/Makefile straight from /root/reference/src/ORBextractor.cc against
// oracle/cvshim).  The resulting oracle/_ref/libref_orb*.so is what tests/ compare the restated oracle and the CUDA path
// against, and what bench.py's cpu_baseline / --impl reference legs time.  Nothing under openmavis_b200/ links or loads it.
//
// The reference's intermediate results are reached the way a C++ user could: `mvImagePyramid` is a public member
// (include/ORBextractor.h:52), ComputePyramid / ComputeKeyPointsOctTree / DistributeOctTree are protected members
// (include/ORBextractor.h:55-64), so a derived class may call them.
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <memory>
#include <thread>
#include <vector>

#include "ORBextractor.h"

namespace {

struct ref_kp { float x, y, size, angle, response; int octave, class_id; };
static_assert(sizeof(cv::KeyPoint) == sizeof(ref_kp), "cv::KeyPoint layout");

class Probe : public ORB_SLAM3::ORBextractor {
  public:
    using ORB_SLAM3::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> out_kps;
    cv::Mat out_desc;
    std::vector<std::vector<cv::KeyPoint>> level_kps;

    void stages(const cv::Mat &img) {   // the first half of operator() (ORBextractor.cc:1001-1004), results kept
        ComputePyramid(img);
        level_kps.clear();
        ComputeKeyPointsOctTree(level_kps);
    }
    std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint> &keys, int minX, int maxX, int minY, int maxY, int N, int level) {
        return DistributeOctTree(keys, minX, maxX, minY, maxY, N, level);
    }
    const std::vector<int> &features_per_level() const { return mnFeaturesPerLevel; }
    const std::vector<int> &u_max() const { return umax; }
};

double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

}  // namespace

extern "C" {

void *ref_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Probe(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_destroy(void *h) { delete (Probe *)h; }

void ref_tables(void *h, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2, int *nfeat, int *umax) {
    Probe *E = (Probe *)h;
    std::vector<float> a = E->GetScaleFactors(), b = E->GetInverseScaleFactors(), c = E->GetScaleSigmaSquares(),
                       d = E->GetInverseScaleSigmaSquares();
    for (int i = 0; i < E->GetLevels(); i++) { sf[i] = a[i]; inv_sf[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; nfeat[i] = E->features_per_level()[i]; }
    for (size_t v = 0; v < E->u_max().size(); v++) umax[v] = E->u_max()[v];
}

// ORBextractor::operator() (ORBextractor.cc:987-1071); returns its return value (number of non-lapping keypoints or -1)
int ref_extract(void *h, const uint8_t *img, int w, int hh, int stride, int lap0, int lap1) {
    Probe *E = (Probe *)h;
    cv::Mat image = (img && w > 0 && hh > 0) ? cv::Mat(hh, w, CV_8UC1, (void *)img, (size_t)stride) : cv::Mat();
    std::vector<int> lap = {lap0, lap1};
    E->out_desc = cv::Mat();
    return (*E)(image, cv::Mat(), E->out_kps, E->out_desc, lap);
}
int ref_num_keypoints(void *h) { return (int)((Probe *)h)->out_kps.size(); }
void ref_result(void *h, ref_kp *kps, uint8_t *desc) {
    Probe *E = (Probe *)h;
    if (E->out_kps.empty()) return;
    memcpy(kps, E->out_kps.data(), E->out_kps.size() * sizeof(ref_kp));
    for (int r = 0; r < E->out_desc.rows; r++) memcpy(desc + (size_t)r * 32, E->out_desc.ptr(r), 32);
}
void ref_level_size(void *h, int l, int *w, int *hh) {
    cv::Mat &m = ((Probe *)h)->mvImagePyramid[l]; *w = m.cols; *hh = m.rows;
}
// the level with its 19-px frame: mvImagePyramid[l] is the ROI at (19,19) of the bordered buffer (ORBextractor.cc:1078-1079)
void ref_level_bordered(void *h, int l, uint8_t *dst) {
    cv::Mat &m = ((Probe *)h)->mvImagePyramid[l];
    const int b = 19, W = m.cols + 2 * b;
    for (int r = 0; r < m.rows + 2 * b; r++) memcpy(dst + (size_t)r * W, m.data + (ptrdiff_t)(r - b) * (ptrdiff_t)m.step - b, W);
}
// ComputePyramid + ComputeKeyPointsOctTree: per-level keypoints (level coordinates, angle set) as the reference holds them
// before the descriptor stage
int ref_stages(void *h, const uint8_t *img, int w, int hh, int stride) {
    Probe *E = (Probe *)h;
    E->stages(cv::Mat(hh, w, CV_8UC1, (void *)img, (size_t)stride));
    int n = 0;
    for (auto &v : E->level_kps) n += (int)v.size();
    return n;
}
int ref_level_num_keypoints(void *h, int l) { return (int)((Probe *)h)->level_kps[l].size(); }
void ref_level_keypoints(void *h, int l, ref_kp *kps) {
    auto &v = ((Probe *)h)->level_kps[l];
    if (!v.empty()) memcpy(kps, v.data(), v.size() * sizeof(ref_kp));
}
// ORBextractor::DistributeOctTree (ORBextractor.cc:496-702) on caller-supplied candidates; out = (x, y, response) per
// selected keypoint in the reference's output order
int ref_distribute(void *h, const int *xs, const int *ys, const int *scores, int n, int minX, int maxX, int minY, int maxY, int N,
                   int *ox, int *oy, int *oscore, int cap) {
    std::vector<cv::KeyPoint> keys(n);
    for (int i = 0; i < n; i++) keys[i] = cv::KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)scores[i]);
    std::vector<cv::KeyPoint> r = ((Probe *)h)->distribute(keys, minX, maxX, minY, maxY, N, 0);
    int m = (int)r.size() < cap ? (int)r.size() : cap;
    for (int i = 0; i < m; i++) { ox[i] = (int)r[i].pt.x; oy[i] = (int)r[i].pt.y; oscore[i] = (int)r[i].response; }
    return (int)r.size();
}

// CPU baseline: nimg images of one geometry through operator(), `nthreads` workers each owning one extractor (the reference
// runs one std::thread per camera, src/Frame.cc:1841-1862).  Returns seconds; total keypoints in *nkp.
double ref_extract_batch(int nimg, const uint8_t *imgs, int w, int hh, int nfeatures, float sf, int nlevels, int iniTh, int minTh,
                         int lap0, int lap1, int nthreads, long long *nkp) {
    if (nthreads < 1) nthreads = 1;
    std::vector<long long> cnt(nthreads, 0);
    std::vector<std::thread> th;
    double t0 = now_s();
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            ORB_SLAM3::ORBextractor E(nfeatures, sf, nlevels, iniTh, minTh);
            std::vector<cv::KeyPoint> kps; cv::Mat desc; std::vector<int> lap = {lap0, lap1};
            for (int i = t; i < nimg; i += nthreads) {
                cv::Mat image(hh, w, CV_8UC1, (void *)(imgs + (size_t)i * w * hh), (size_t)w);
                E(image, cv::Mat(), kps, desc, lap);
                cnt[t] += (long long)kps.size();
            }
        });
    for (auto &x : th) x.join();
    double dt = now_s() - t0;
    long long s = 0;
    for (auto c : cnt) s += c;
    if (nkp) *nkp = s;
    return dt;
}

// CPU baseline, reference-faithful threading: rig frames one after the other, each with one std::thread per camera created and joined
// per frame around persistent extractor objects (src/Frame.cc:1841-1862 + mpORBextractorLeft ... of src/Tracking.cc:572-587).
// imgs = [nframes][ncam][h][w]; laps = [ncam][2].  Returns seconds.
double ref_extract_rig_frames(int nframes, int ncam, const uint8_t *imgs, int w, int hh, int nfeatures, float sf, int nlevels, int iniTh, int minTh,
                              const int *laps, long long *nkp) {
    std::vector<std::unique_ptr<ORB_SLAM3::ORBextractor>> ex;
    for (int c = 0; c < ncam; c++) ex.emplace_back(new ORB_SLAM3::ORBextractor(nfeatures, sf, nlevels, iniTh, minTh));
    std::vector<std::vector<cv::KeyPoint>> kps(ncam);
    std::vector<cv::Mat> desc(ncam);
    long long total = 0;
    const double t0 = now_s();
    for (int f = 0; f < nframes; f++) {
        std::vector<std::thread> th;
        for (int c = 0; c < ncam; c++)
            th.emplace_back([&, c]() {
                cv::Mat image(hh, w, CV_8UC1, (void *)(imgs + ((size_t)f * ncam + c) * w * hh), (size_t)w);
                std::vector<int> lap = {laps[2 * c], laps[2 * c + 1]};
                (*ex[c])(image, cv::Mat(), kps[c], desc[c], lap);
            });
        for (auto &t : th) t.join();
        for (int c = 0; c < ncam; c++) total += (long long)kps[c].size();
    }
    const double dt = now_s() - t0;
    if (nkp) *nkp = total;
    return dt;
}

}  // extern "C"
