// ORBextractor.cc — ORB_SLAM3::ORBextractor on top of liborbx.so (replaces src/ORBextractor.cc of the reference).
#include "ORBextractor.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>

#include "../../../include/orbx.h"

namespace ORB_SLAM3 {

// ------------------------------------------------------------------------------------------------------------------------------
// Rig rendezvous.  The reference calls operator() from 2-4 std::threads, one per camera / extractor object, and joins them right
// away (src/Frame.cc:211-214, 1320-1337, 1841-1862).  Extractors of the same geometry on the same device share ONE batch-capable
// device handle: the first caller of a rig frame becomes the leader, waits a few hundred microseconds for the calls that usually
// arrive with it, and launches one orbx_extract_batch for all of them (per-image nfeatures / thresholds / lapping area); the
// others sleep until their slot is filled.  "Usually arrive with it" is learned: every batch records how many calls it took plus how
// many arrived while it ran, the maximum over the last four observations is what the next leader waits for — so a monocular
// pipeline (one caller) never waits, and Frame.cc needs no change.  ORBX_RIG_WAIT_US sets the timeout (default 500, 0 = never wait).
// ------------------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kRigMax = 8;

struct RigSlot {
    const unsigned char *img; int stride;
    orbx_image_params prm;
    orbx_kp *kps; unsigned char *desc; int cap;
    std::vector<cv::Mat> *pyramid;      // non-null: download the levels of this image
    int n = 0, mono = 0, status = 0;
    bool done = false;
};

struct RigContext {
    std::mutex mu;
    std::condition_variable cv;
    orbx_extractor *handle = nullptr;
    int device = 0, w = 0, h = 0, nlevels = 0; float sf = 0;
    int handle_images = 0, handle_nfeat = 0;
    std::vector<RigSlot *> pending;
    bool busy = false;
    int hist[4] = {1, 1, 1, 1}, hist_pos = 0;
    long long batches = 0, images = 0;
    ~RigContext() { orbx_extractor_destroy(handle); }
    int expected() const { return std::max(std::max(hist[0], hist[1]), std::max(hist[2], hist[3])); }
};

std::mutex g_rig_mu;
std::map<std::tuple<int, int, int, int, float>, std::shared_ptr<RigContext>> g_rig;
long long g_batches = 0, g_images = 0;

std::shared_ptr<RigContext> rig_context(int device, int w, int h, int nlevels, float sf) {
    std::lock_guard<std::mutex> lk(g_rig_mu);
    std::shared_ptr<RigContext> &c = g_rig[std::make_tuple(device, w, h, nlevels, sf)];
    if (!c) { c.reset(new RigContext()); c->device = device; c->w = w; c->h = h; c->nlevels = nlevels; c->sf = sf; }
    return c;
}

int rig_wait_us() {
    static const int us = [] { const char *t = getenv("ORBX_RIG_WAIT_US"); return t ? std::max(0, atoi(t)) : 500; }();
    return us;
}

[[noreturn]] void die(const char *what) {
    fprintf(stderr, "ORBextractor: %s: %s\n", what, orbx_last_error());   // no CPU fallback: fail loudly
    abort();
}

// runs one batch on the shared handle (called by the leader, handle owned exclusively while ctx.busy)
void rig_run(RigContext &C, std::vector<RigSlot *> &batch) {
    const int n = (int)batch.size();
    int max_nf = 1;
    for (RigSlot *s : batch) max_nf = std::max(max_nf, s->prm.nfeatures);
    if (!C.handle || C.handle_images < n || C.handle_nfeat < max_nf) {
        orbx_extractor_destroy(C.handle);
        C.handle = nullptr;
        C.handle_images = std::max(std::max(C.handle_images, n), 4); C.handle_nfeat = std::max(C.handle_nfeat, max_nf);
        if (orbx_extractor_create(&C.handle, C.device, C.w, C.h, C.handle_images, C.handle_nfeat, C.sf, C.nlevels, 0) != ORBX_OK) die("orbx_extractor_create");
    }
    const int cap = orbx_extractor_capacity(C.handle);
    std::vector<const unsigned char *> imgs(n);
    std::vector<int> strides(n), nk(n), nm(n);
    std::vector<orbx_image_params> prm(n);
    std::vector<orbx_kp> kps((size_t)n * cap);
    std::vector<unsigned char> desc((size_t)n * cap * 32);
    for (int i = 0; i < n; i++) { imgs[i] = batch[i]->img; strides[i] = batch[i]->stride; prm[i] = batch[i]->prm; }
    if (orbx_extract_batch(C.handle, n, imgs.data(), strides.data(), prm.data(), kps.data(), desc.data(), cap, nk.data(), nm.data()) != ORBX_OK) die("orbx_extract_batch");
    std::vector<int> lw(C.nlevels), lh(C.nlevels);
    orbx_extractor_tables(C.handle, nullptr, nullptr, nullptr, nullptr, lw.data(), lh.data());
    for (int i = 0; i < n; i++) {
        RigSlot &s = *batch[i];
        s.n = nk[i]; s.mono = nm[i];
        if (s.n > s.cap) die("internal: slot capacity");
        memcpy(s.kps, &kps[(size_t)i * cap], (size_t)s.n * sizeof(orbx_kp));
        memcpy(s.desc, &desc[(size_t)i * cap * 32], (size_t)s.n * 32);
        if (s.pyramid)
            for (int l = 0; l < C.nlevels; l++) {
                // like the reference, a level is the interior ROI of a (w+38) x (h+38) buffer with the REFLECT_101 frame
                cv::Mat temp(lh[l] + 38, lw[l] + 38, CV_8UC1);
                orbx_extractor_pyramid_level(C.handle, i, l, 1, temp.data, (int)temp.step);
#ifdef ORBX_HAVE_OPENCV
                (*s.pyramid)[l] = temp(cv::Rect(19, 19, lw[l], lh[l]));
#else
                (*s.pyramid)[l] = temp.roi(19, 19, lw[l], lh[l]);
#endif
            }
    }
    C.batches++; C.images += n;
    { std::lock_guard<std::mutex> lk(g_rig_mu); g_batches++; g_images += n; }
}

void rig_submit(RigContext &C, RigSlot &s) {
    std::unique_lock<std::mutex> lk(C.mu);
    C.pending.push_back(&s);
    if (C.pending.size() > 1) {                         // follower: the leader of this batch fills the slot
        if ((int)C.pending.size() >= C.expected()) C.cv.notify_all();
        C.cv.wait(lk, [&] { return s.done; });
        return;
    }
    // leader
    const int want = std::min(C.expected(), kRigMax);
    if (want > 1 && rig_wait_us() > 0)
        C.cv.wait_for(lk, std::chrono::microseconds(rig_wait_us()), [&] { return (int)C.pending.size() >= want; });
    C.cv.wait(lk, [&] { return !C.busy; });             // the previous batch still owns the handle: later callers keep joining
    std::vector<RigSlot *> batch;
    const int take = std::min((int)C.pending.size(), kRigMax);
    batch.assign(C.pending.begin(), C.pending.begin() + take);
    C.pending.erase(C.pending.begin(), C.pending.begin() + take);
    C.busy = true;
    if (!C.pending.empty()) C.cv.notify_all();          // (more than kRigMax callers: the next one leads its own batch)
    lk.unlock();
    rig_run(C, batch);
    lk.lock();
    C.busy = false;
    C.hist[C.hist_pos++ & 3] = std::min(kRigMax, (int)batch.size() + (int)C.pending.size());
    for (RigSlot *b : batch) b->done = true;
    C.cv.notify_all();
}

}  // namespace

// number of device batches / images the rendezvous has launched in this process (tests, diagnostics)
extern "C" void orbx_shim_rig_stats(long long *batches, long long *images) {
    std::lock_guard<std::mutex> lk(g_rig_mu);
    if (batches) *batches = g_batches;
    if (images) *images = g_images;
}

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

ORBextractor::~ORBextractor() {}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                             cv::OutputArray _descriptors, std::vector<int> &vLappingArea) {
    if (_image.empty()) return -1;                       // src/ORBextractor.cc:994-995
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) { fprintf(stderr, "ORBextractor: image.type() must be CV_8UC1\n"); abort(); }   // :998 assert
    if (!mRig || image.cols != mWidth || image.rows != mHeight) {
        mRig = rig_context(mDevice, image.cols, image.rows, nlevels, (float)scaleFactor);
        mWidth = image.cols; mHeight = image.rows;
    }
    RigContext &C = *std::static_pointer_cast<RigContext>(mRig);
    const int cap = nfeatures + 16 * nlevels + 64;       // the quadtree overshoots a level's target by at most 2: ample
    mKpBuf.resize((size_t)cap * sizeof(orbx_kp));
    mDescBuf.resize((size_t)cap * 32);
    RigSlot s;
    s.img = image.data; s.stride = (int)image.step;
    s.prm.nfeatures = nfeatures; s.prm.ini_th_fast = iniThFAST; s.prm.min_th_fast = minThFAST;
    s.prm.lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0;
    s.prm.lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
    s.kps = (orbx_kp *)mKpBuf.data(); s.desc = mDescBuf.data(); s.cap = cap;
    s.pyramid = mbDownloadPyramid ? &mvImagePyramid : nullptr;
    rig_submit(C, s);
    const int n = s.n, mono = s.mono;
    // callee allocates into the caller's containers (src/ORBextractor.cc:1013-1022)
    _keypoints = std::vector<cv::KeyPoint>(n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint layout");
    if (n) memcpy((void *)_keypoints.data(), mKpBuf.data(), (size_t)n * sizeof(orbx_kp));
    _descriptors.create(n, 32, CV_8U);
    cv::Mat descriptors = _descriptors.getMat();
    for (int i = 0; i < n; i++) memcpy(descriptors.ptr<unsigned char>(i), mDescBuf.data() + (size_t)i * 32, 32);
    return mono;
}

}  // namespace ORB_SLAM3
