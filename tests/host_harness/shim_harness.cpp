// C entry points around the C++ shim classes (openmavis_b200/csrc/host/) so the Python test-suite can drive
// ORB_SLAM3::ORBextractor / ORBmatcher exactly the way src/Frame.cc does.  Test infrastructure.
#include <string.h>

#include <thread>
#include <vector>

#include "../../openmavis_b200/csrc/host/ORBextractor.h"
#include "../../openmavis_b200/csrc/host/ORBmatcher.h"

extern "C" {

void *shim_extractor_new(int nf, float sf, int nl, int ini, int mn) { return new ORB_SLAM3::ORBextractor(nf, sf, nl, ini, mn); }
void shim_extractor_delete(void *e) { delete (ORB_SLAM3::ORBextractor *)e; }

void shim_extractor_tables(void *e, float *sf, float *inv_sf, float *s2, float *inv_s2) {
    ORB_SLAM3::ORBextractor *x = (ORB_SLAM3::ORBextractor *)e;
    const int n = x->GetLevels();
    std::vector<float> a = x->GetScaleFactors(), b = x->GetInverseScaleFactors(), c = x->GetScaleSigmaSquares(), d = x->GetInverseScaleSigmaSquares();
    memcpy(sf, a.data(), n * 4); memcpy(inv_sf, b.data(), n * 4); memcpy(s2, c.data(), n * 4); memcpy(inv_s2, d.data(), n * 4);
}

// the call of Frame::ExtractORB (src/Frame.cc:599-640): returns monoIndex; *n = keypoints.size()
int shim_extract(void *e, unsigned char *img, int w, int h, int stride, int lap0, int lap1, float *kps28, unsigned char *desc, int cap, int *n) {
    ORB_SLAM3::ORBextractor *x = (ORB_SLAM3::ORBextractor *)e;
    cv::Mat image = img ? cv::Mat(h, w, CV_8UC1, img, (size_t)stride) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    std::vector<int> lap = {lap0, lap1};
    const int mono = (*x)(image, cv::Mat(), keys, descriptors, lap);
    *n = (int)keys.size();
    const int m = *n < cap ? *n : cap;
    if (m > 0) memcpy(kps28, keys.data(), (size_t)m * sizeof(cv::KeyPoint));
    for (int i = 0; i < m; i++) memcpy(desc + (size_t)i * 32, descriptors.ptr<unsigned char>(i), 32);
    return mono;
}

void shim_set_pyramid_download(void *e, int on) { ((ORB_SLAM3::ORBextractor *)e)->SetPyramidDownload(on != 0); }

extern "C" void orbx_shim_rig_stats(long long *batches, long long *images);
void shim_rig_stats(long long *batches, long long *images) { orbx_shim_rig_stats(batches, images); }

// The multi-camera Frame constructor (src/Frame.cc:1841-1862): per rig frame one std::thread per camera, each calling operator() of its
// own extractor object, joined right away.  imgs = [frames][ncam][h][w]; nf / laps per camera; outputs of the LAST frame:
// kps28 [ncam][cap], desc [ncam][cap][32], n [ncam], mono [ncam].
void shim_extract_rig(int ncam, int frames, unsigned char *imgs, int w, int h, const int *nf, const int *laps, int ini, int mn, float *kps28,
                      unsigned char *desc, int cap, int *n, int *mono) {
    std::vector<ORB_SLAM3::ORBextractor *> ex;
    for (int c = 0; c < ncam; c++) ex.push_back(new ORB_SLAM3::ORBextractor(nf[c], 1.2f, 8, ini, mn));
    std::vector<std::vector<cv::KeyPoint>> keys(ncam);
    std::vector<cv::Mat> descs(ncam);
    for (int f = 0; f < frames; f++) {
        std::vector<std::thread> th;
        for (int c = 0; c < ncam; c++)
            th.emplace_back([&, c]() {
                cv::Mat image(h, w, CV_8UC1, imgs + ((size_t)f * ncam + c) * w * h, (size_t)w);
                std::vector<int> lap = {laps[2 * c], laps[2 * c + 1]};
                mono[c] = (*ex[c])(image, cv::Mat(), keys[c], descs[c], lap);
            });
        for (auto &t : th) t.join();
    }
    for (int c = 0; c < ncam; c++) {
        n[c] = (int)keys[c].size();
        const int m = n[c] < cap ? n[c] : cap;
        if (m > 0) memcpy(kps28 + (size_t)c * cap * 7, keys[c].data(), (size_t)m * sizeof(cv::KeyPoint));
        for (int i = 0; i < m; i++) memcpy(desc + ((size_t)c * cap + i) * 32, descs[c].ptr<unsigned char>(i), 32);
        delete ex[c];
    }
}

// copies mvImagePyramid[level] (interior) into dst (tightly packed); returns 0 if the level is empty
int shim_pyramid_level(void *e, int level, unsigned char *dst, int *w, int *h) {
    ORB_SLAM3::ORBextractor *x = (ORB_SLAM3::ORBextractor *)e;
    const cv::Mat &m = x->mvImagePyramid[level];
    if (m.empty()) return 0;
    *w = m.cols; *h = m.rows;
    if (dst) for (int r = 0; r < m.rows; r++) memcpy(dst + (size_t)r * m.cols, m.ptr<unsigned char>(r), m.cols);
    return 1;
}

int shim_descriptor_distance(unsigned char *a, unsigned char *b) {
    return ORB_SLAM3::ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, a, 32), cv::Mat(1, 32, CV_8U, b, 32));
}

int shim_knn_ratio(unsigned char *q, int nq, unsigned char *t, int nt, double ratio, int *idx, int *dist, unsigned char *keep) {
    ORB_SLAM3::ORBmatcher m(0.8f, false);
    std::vector<int> vi, vd;
    m.KnnMatch2(cv::Mat(nq, 32, CV_8U, q, 32), cv::Mat(nt, 32, CV_8U, t, 32), vi, vd);
    std::vector<unsigned char> k = ORB_SLAM3::ORBmatcher::RatioTest(vd, ratio);
    memcpy(idx, vi.data(), vi.size() * 4); memcpy(dist, vd.data(), vd.size() * 4); memcpy(keep, k.data(), k.size());
    return ORB_SLAM3::ORBmatcher::TH_LOW * 1000 + ORB_SLAM3::ORBmatcher::TH_HIGH;
}

// SearchForInitialization through the shim class (std::vector<cv::KeyPoint> / cv::Point2f in, vnMatches12 out)
int shim_search_for_initialization(float *kps1_28, unsigned char *d1, int n1, float *kps2_28, unsigned char *d2, int n2, float *prev_xy,
                                   int window, float nnratio, int check, int *m12) {
    ORB_SLAM3::ORBmatcher matcher(nnratio, check != 0);
    std::vector<cv::KeyPoint> k1(n1), k2(n2);
    memcpy((void *)k1.data(), kps1_28, (size_t)n1 * 28); memcpy((void *)k2.data(), kps2_28, (size_t)n2 * 28);
    cv::Mat m1(n1, 32, CV_8U, d1, 32), m2(n2, 32, CV_8U, d2, 32);
    std::vector<cv::Point2f> prev(n1);
    memcpy((void *)prev.data(), prev_xy, (size_t)n1 * 8);
    std::vector<int> v12;
    const int nm = matcher.SearchForInitialization(k1, m1, k2, m2, 0.f, 512.f, 0.f, 512.f, prev, v12, window);
    memcpy(m12, v12.data(), (size_t)n1 * 4); memcpy(prev_xy, prev.data(), (size_t)n1 * 8);
    return nm;
}

}  // extern "C"
