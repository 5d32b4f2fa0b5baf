// orb_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// An OpenCV-free C++ restatement of the reference's ORB extraction / Hamming matching hot path
// (OpenMAVIS, an ORB-SLAM3 fork; all file:line cites are relative to /root/reference).  It exists to
// CHECK the CUDA product in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs.  Nothing under openmavis_b200/ may include, link or call it.
//
// Parity status: the reference ships no golden vectors for this path, and its own
// sources cannot be compiled here (no OpenCV/Eigen/Sophus headers).  The oracle is therefore pinned
// primitive by primitive against the reference's third-party dependency itself — OpenCV, through the
// cv2 4.13.0 wheel in this image (tests/test_oracle_vs_cv2.py: resize, copyMakeBorder, FAST, GaussianBlur,
// fastAtan2, BFMatcher.knnMatch) — and against committed fixtures generated from cv2
// (tests/golden/, tools/gen_golden.py).  The glue between primitives (cell grid, quadtree
// distribution, orientation, rBRIEF, lapping-area partition, matcher bookkeeping) follows the
// reference line by line and is additionally checked by an independent Python/cv2 restatement.
//
// Build: oracle/Makefile  (g++ -O3 -march=native -ffp-contract=off, mirrors CMakeLists.txt:13-16
// except that contraction is off; the FMA behaviour of GET_VALUE is an explicit switch, see below).

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>

#include "../include/orbx_pattern.h"

namespace {

const int kPatchSize = 31;       // ORBextractor.cc:15
const int kHalfPatch = 15;       // ORBextractor.cc:16
const int kEdge = 19;            // ORBextractor.cc:17 (EDGE_THRESHOLD)
const signed char kPattern[1024] = {ORBX_PATTERN_VALUES};

inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound(float): cvtss2si, half-to-even
inline int cv_round_d(double v) { return (int)lrint(v); }   // cvRound(double)

struct KeyPt {   // mirrors cv::KeyPoint field for field
    float x, y, size, angle, response;
    int octave, class_id;
};

// ----------------------------------------------------------------------------------------------
// Per-level tables.  ORBextractor.cc:351-414.
// ----------------------------------------------------------------------------------------------
struct Params {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;   // double member holding the float argument (ORBextractor.h:70)
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
    std::vector<int> nfeat;
    int umax[kHalfPatch + 1];
};

void make_params(Params &p, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    p.nfeatures = nfeatures; p.nlevels = nlevels; p.iniTh = iniTh; p.minTh = minTh;
    p.scaleFactor = scaleFactor;
    p.sf.assign(nlevels, 1.0f); p.sigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; i++) {
        p.sf[i] = (float)(p.sf[i - 1] * p.scaleFactor);          // :366
        p.sigma2[i] = p.sf[i] * p.sf[i];                         // :367
    }
    p.inv_sf.resize(nlevels); p.inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        p.inv_sf[i] = 1.0f / p.sf[i];
        p.inv_sigma2[i] = 1.0f / p.sigma2[i];
    }
    p.nfeat.assign(nlevels, 0);
    float factor = (float)(1.0f / p.scaleFactor);                // :380
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        p.nfeat[l] = cv_round_f(nDesired);
        sum += p.nfeat[l];
        nDesired *= factor;
    }
    p.nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax: :399-413
    int v, v0;
    int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= kHalfPatch; v++) p.umax[v] = 0;
    for (v = 0; v <= vmax; ++v) p.umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (p.umax[v0] == p.umax[v0 + 1]) ++v0;
        p.umax[v] = v0;
        ++v0;
    }
}

// ----------------------------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) for CV_8UC1 — fixed-point restatement (OpenCV imgproc/resize.cpp,
// HResizeLinear / VResizeLinear with INTER_RESIZE_COEF_BITS = 11).  Called at ORBextractor.cc:1083.
// ----------------------------------------------------------------------------------------------
void linear_taps(int src, int dst, std::vector<int> &ofs, std::vector<short> &c0, std::vector<short> &c1) {
    ofs.resize(dst); c0.resize(dst); c1.resize(dst);
    double inv_scale = (double)dst / src;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        ofs[d] = s;
        c0[d] = (short)cv_round_f((1.f - f) * 2048.f);
        c1[d] = (short)cv_round_f(f * 2048.f);
    }
}

void resize_linear(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride) {
    std::vector<int> xo, yo; std::vector<short> a0, a1, b0, b1;
    linear_taps(sw, dw, xo, a0, a1);
    linear_taps(sh, dh, yo, b0, b1);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; y++) {
        int sy0 = yo[y], sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t *S0 = src + (size_t)sy0 * sstride, *S1 = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; x++) {
            int sx0 = xo[x], sx1 = std::min(sx0 + 1, sw - 1);
            r0[x] = S0[sx0] * a0[x] + S0[sx1] * a1[x];
            r1[x] = S1[sx0] * a0[x] + S1[sx1] * a1[x];
        }
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++)
            D[x] = (uint8_t)((((b0[y] * (r0[x] >> 4)) >> 16) + ((b1[y] * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// cv::copyMakeBorder(BORDER_REFLECT_101) around a w x h interior that already sits at (b,b) of buf.
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}
void fill_border_reflect101(uint8_t *buf, int w, int h, int stride, int b) {
    uint8_t *in = buf + (size_t)b * stride + b;
    for (int y = 0; y < h; y++) {
        uint8_t *row = in + (size_t)y * stride;
        for (int k = 1; k <= b; k++) { row[-k] = row[reflect101(-k, w)]; row[w - 1 + k] = row[reflect101(w - 1 + k, w)]; }
    }
    for (int k = 1; k <= b; k++) {
        memcpy(buf + (size_t)(b - k) * stride, buf + (size_t)(b + reflect101(-k, h)) * stride, w + 2 * b);
        memcpy(buf + (size_t)(b + h - 1 + k) * stride, buf + (size_t)(b + reflect101(h - 1 + k, h)) * stride, w + 2 * b);
    }
}

// ----------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression=true)  (TYPE_9_16) — restatement of OpenCV
// features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>.  Called at ORBextractor.cc:745,765.
// ----------------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int corner_score16(const uint8_t *p, const int *off, int threshold) {
    const int K = 8, N = K * 3 + 1;
    int v = p[0];
    short d[N];
    for (int k = 0; k < N; k++) d[k] = (short)(v - p[off[k]]);
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min((int)d[k + 1], (int)d[k + 2]);
        a = std::min(a, (int)d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, (int)d[k + 4]); a = std::min(a, (int)d[k + 5]);
        a = std::min(a, (int)d[k + 6]); a = std::min(a, (int)d[k + 7]);
        a = std::min(a, (int)d[k + 8]);
        a0 = std::max(a0, std::min(a, (int)d[k]));
        a0 = std::max(a0, std::min(a, (int)d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max((int)d[k + 1], (int)d[k + 2]);
        b = std::max(b, (int)d[k + 3]); b = std::max(b, (int)d[k + 4]);
        b = std::max(b, (int)d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, (int)d[k + 6]); b = std::max(b, (int)d[k + 7]);
        b = std::max(b, (int)d[k + 8]);
        b0 = std::min(b0, std::max(b, (int)d[k]));
        b0 = std::min(b0, std::max(b, (int)d[k + 9]));
    }
    return -b0 - 1;
}

struct FastPt { int x, y, score; };

void fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, std::vector<FastPt> &out) {
    out.clear();
    if (w < 7 || h < 7) return;
    threshold = std::min(std::max(threshold, 0), 255);
    int off[25];
    for (int k = 0; k < 25; k++) off[k] = kRingDy[k % 16] * stride + kRingDx[k % 16];
    uint8_t ttab[512];
    for (int i = -255; i <= 255; i++) ttab[i + 255] = (uint8_t)(i < -threshold ? 1 : i > threshold ? 2 : 0);
    std::vector<uint8_t> sbuf((size_t)3 * w, 0);
    std::vector<int> cbuf((size_t)3 * (w + 1), 0);
    uint8_t *buf[3] = {sbuf.data(), sbuf.data() + w, sbuf.data() + 2 * w};
    int *cp[3] = {cbuf.data(), cbuf.data() + (w + 1), cbuf.data() + 2 * (w + 1)};
    for (int i = 3; i < h - 2; i++) {
        const uint8_t *ptr = img + (size_t)i * stride + 3;
        uint8_t *curr = buf[(i - 3) % 3];
        int *cornerpos = cp[(i - 3) % 3] + 1;
        memset(curr, 0, w);
        int ncorners = 0;
        if (i < h - 3) {
            for (int j = 3; j < w - 3; j++, ptr++) {
                int v = ptr[0];
                // high-speed test of fast.cpp (a necessary condition for an arc of 9): tab = 1 darker, 2 brighter
                const uint8_t *tab = &ttab[0] - v + 255;
                int d = tab[ptr[off[0]]] | tab[ptr[off[8]]];
                if (d == 0) continue;
                d &= tab[ptr[off[2]]] | tab[ptr[off[10]]];
                d &= tab[ptr[off[4]]] | tab[ptr[off[12]]];
                d &= tab[ptr[off[6]]] | tab[ptr[off[14]]];
                if (d == 0) continue;
                d &= tab[ptr[off[1]]] | tab[ptr[off[9]]];
                d &= tab[ptr[off[3]]] | tab[ptr[off[11]]];
                d &= tab[ptr[off[5]]] | tab[ptr[off[13]]];
                d &= tab[ptr[off[7]]] | tab[ptr[off[15]]];
                int lo = v - threshold, hi = v + threshold;
                // >= 9 contiguous ring pixels all darker than v-t, or all brighter than v+t
                bool corner = false;
                if (d & 1) {
                    int count = 0;
                    for (int k = 0; k < 25; k++) {
                        if (ptr[off[k]] < lo) { if (++count > 8) { corner = true; break; } } else count = 0;
                    }
                }
                if (!corner && (d & 2)) {
                    int count = 0;
                    for (int k = 0; k < 25; k++) {
                        if (ptr[off[k]] > hi) { if (++count > 8) { corner = true; break; } } else count = 0;
                    }
                }
                if (corner) {
                    cornerpos[ncorners++] = j;
                    curr[j] = (uint8_t)corner_score16(ptr, off, threshold);
                }
            }
        }
        cornerpos[-1] = ncorners;
        if (i == 3) continue;
        const uint8_t *prev = buf[(i - 4 + 3) % 3];
        const uint8_t *pprev = buf[(i - 5 + 3) % 3];
        cornerpos = cp[(i - 4 + 3) % 3] + 1;
        ncorners = cornerpos[-1];
        for (int k = 0; k < ncorners; k++) {
            int j = cornerpos[k];
            int score = prev[j];
            if (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] && score > pprev[j] &&
                score > pprev[j + 1] && score > curr[j - 1] && score > curr[j] && score > curr[j + 1])
                out.push_back({j, i - 1, score});
        }
    }
}

// ----------------------------------------------------------------------------------------------
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) for CV_8U — OpenCV's fixed-point (Q8.8) path:
// kernel {18,34,48,56,48,34,18}/256, horizontal u8->u16, vertical u16->u32, (v + 2^15) >> 16.
// Called at ORBextractor.cc:1036 on a borderless clone.
// ----------------------------------------------------------------------------------------------
const int kBlurK[7] = {18, 34, 48, 56, 48, 34, 18};
void blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    std::vector<uint16_t> tmp((size_t)w * h);
    std::vector<uint8_t> row(w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstride;
        for (int k = 0; k < 3; k++) { row[k] = S[reflect101(k - 3, w)]; row[w + 3 + k] = S[reflect101(w + k, w)]; }
        memcpy(row.data() + 3, S, w);
        uint16_t *T = tmp.data() + (size_t)y * w;
        const uint8_t *R = row.data();
        for (int x = 0; x < w; x++)
            T[x] = (uint16_t)(18 * (R[x] + R[x + 6]) + 34 * (R[x + 1] + R[x + 5]) + 48 * (R[x + 2] + R[x + 4]) + 56 * R[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *R[7];
        for (int k = 0; k < 7; k++) R[k] = tmp.data() + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 18u * ((uint32_t)R[0][x] + R[6][x]) + 34u * ((uint32_t)R[1][x] + R[5][x]) +
                           48u * ((uint32_t)R[2][x] + R[4][x]) + 56u * (uint32_t)R[3][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// cv::fastAtan2(y, x) in degrees — OpenCV core/mathfuncs_core (scalar atan_f32), each op fp32-rounded.
float fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps); c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps); c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// IC_Angle.  ORBextractor.cc:19-43.
float ic_angle(const uint8_t *img, int stride, int x, int y, const int *umax) {
    int m01 = 0, m10 = 0;
    const uint8_t *c = img + (size_t)y * stride + x;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = c[u + v * stride], vm = c[u - v * stride];
            vsum += (vp - vm);
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor.  ORBextractor.cc:46-90.  fma_mode 0: every op rounded separately;
// fma_mode 1: GET_VALUE contracted the way g++ -O3 -march=<fma cpu> does it
// (fma(px, b, py*a), fma(px, a, -(py*b))) — SURVEY.md §8c "oracle definition".
void orb_descriptor(const uint8_t *img, int stride, int x, int y, float angle_deg, int fma_mode, uint8_t *desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = angle_deg * factorPI;
    float a = (float)cosf(angle), b = (float)sinf(angle);
    const uint8_t *c = img + (size_t)y * stride + x;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t[2];
            for (int e = 0; e < 2; e++) {
                float px = kPattern[(i * 16 + 2 * k + e) * 2], py = kPattern[(i * 16 + 2 * k + e) * 2 + 1];
                float fy, fx;
                if (fma_mode) { fy = fmaf(px, b, py * a); fx = fmaf(px, a, -(py * b)); }
                else { fy = px * b + py * a; fx = px * a - py * b; }
                t[e] = c[cv_round_f(fy) * stride + cv_round_f(fx)];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ----------------------------------------------------------------------------------------------
// DistributeOctTree.  ORBextractor.cc:425-702 — std::list / std::sort faithful restatement.
// Keys are indices into the candidate array (the reference copies KeyPoints; order is what matters).
// ----------------------------------------------------------------------------------------------
struct Node {
    std::vector<int> keys;
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    std::list<Node>::iterator lit;
    bool noMore = false;
};

void divide_node(const Node &n, const std::vector<FastPt> &pts, Node &n1, Node &n2, Node &n3, Node &n4) {
    const int halfX = (int)ceilf((float)(n.urx - n.ulx) / 2);
    const int halfY = (int)ceilf((float)(n.bry - n.uly) / 2);
    n1.ulx = n.ulx; n1.uly = n.uly; n1.urx = n.ulx + halfX; n1.ury = n.uly;
    n1.blx = n.ulx; n1.bly = n.uly + halfY; n1.brx = n.ulx + halfX; n1.bry = n.uly + halfY;
    n2.ulx = n1.urx; n2.uly = n1.ury; n2.urx = n.urx; n2.ury = n.ury;
    n2.blx = n1.brx; n2.bly = n1.bry; n2.brx = n.urx; n2.bry = n.uly + halfY;
    n3.ulx = n1.blx; n3.uly = n1.bly; n3.urx = n1.brx; n3.ury = n1.bry;
    n3.blx = n.blx; n3.bly = n.bly; n3.brx = n1.brx; n3.bry = n.bly;
    n4.ulx = n3.urx; n4.uly = n3.ury; n4.urx = n2.brx; n4.ury = n2.bry;
    n4.blx = n3.brx; n4.bly = n3.bry; n4.brx = n.brx; n4.bry = n.bry;
    for (int k : n.keys) {
        float px = (float)pts[k].x, py = (float)pts[k].y;
        if (px < n1.urx) { if (py < n1.bry) n1.keys.push_back(k); else n3.keys.push_back(k); }
        else if (py < n1.bry) n2.keys.push_back(k);
        else n4.keys.push_back(k);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

bool compare_nodes(std::pair<int, Node *> &e1, std::pair<int, Node *> &e2) {   // :482-494
    if (e1.first < e2.first) return true;
    if (e1.first > e2.first) return false;
    return e1.second->ulx < e2.second->ulx;
}

// returns the selected candidate indices, in std::list order
void distribute_octtree(const std::vector<FastPt> &pts, int minX, int maxX, int minY, int maxY, int N,
                        std::vector<int> &result) {
    result.clear();
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node *> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ulx = (int)(hX * (float)i); ni.uly = 0;
        ni.urx = (int)(hX * (float)(i + 1)); ni.ury = 0;
        ni.blx = ni.ulx; ni.bly = maxY - minY;
        ni.brx = ni.urx; ni.bry = maxY - minY;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < pts.size(); i++) ini[(size_t)((float)pts[i].x / hX)]->keys.push_back((int)i);
    auto lit = nodes.begin();
    while (lit != nodes.end()) {
        if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
        else if (lit->keys.empty()) lit = nodes.erase(lit);
        else lit++;
    }
    bool finish = false;
    std::vector<std::pair<int, Node *>> sizeAndNode;
    auto push_children = [&](Node *ch[4], int *nToExpand) {
        for (int c = 0; c < 4; c++) {
            if (ch[c]->keys.size() > 0) {
                nodes.push_front(*ch[c]);
                if (ch[c]->keys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    sizeAndNode.push_back(std::make_pair((int)ch[c]->keys.size(), &nodes.front()));
                    nodes.front().lit = nodes.begin();
                }
            }
        }
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        lit = nodes.begin();
        int nToExpand = 0;
        sizeAndNode.clear();
        while (lit != nodes.end()) {
            if (lit->noMore) { lit++; continue; }
            Node n1, n2, n3, n4; Node *ch[4] = {&n1, &n2, &n3, &n4};
            divide_node(*lit, pts, n1, n2, n3, n4);
            push_children(ch, &nToExpand);
            lit = nodes.erase(lit);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if (((int)nodes.size() + nToExpand * 3) > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                std::vector<std::pair<int, Node *>> prev = sizeAndNode;
                sizeAndNode.clear();
                std::sort(prev.begin(), prev.end(), compare_nodes);
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4; Node *ch[4] = {&n1, &n2, &n3, &n4};
                    divide_node(*prev[j].second, pts, n1, n2, n3, n4);
                    push_children(ch, nullptr);
                    nodes.erase(prev[j].second->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    for (auto &n : nodes) {   // :683-699 — first strictly-greatest response
        int best = n.keys[0];
        float maxResp = (float)pts[best].score;
        for (size_t k = 1; k < n.keys.size(); k++)
            if ((float)pts[n.keys[k]].score > maxResp) { best = n.keys[k]; maxResp = (float)pts[best].score; }
        result.push_back(best);
    }
}

// ----------------------------------------------------------------------------------------------
// The extractor object: ComputePyramid (:1073-1104), ComputeKeyPointsOctTree (:704-823),
// operator() (:987-1071).  Keeps every stage's output so tests can compare stage by stage.
// ----------------------------------------------------------------------------------------------
struct Level {
    int w = 0, h = 0, stride = 0;        // interior size, bordered-buffer stride
    std::vector<uint8_t> buf;            // (h+38) x stride, interior at (19,19)
    std::vector<uint8_t> blur;           // w x h borderless (only when the level has keypoints)
    std::vector<FastPt> cand;            // vToDistributeKeys (border-relative coords)
    std::vector<KeyPt> kps;              // after DistributeOctTree + orientation (level coords)
    bool blurred = false;
};

struct Extractor {
    Params p;
    std::vector<Level> lv;
    std::vector<KeyPt> out_kps;
    std::vector<uint8_t> out_desc;
    int n_mono = 0;
    double t_stage[6] = {0, 0, 0, 0, 0, 0};   // pyramid, fast, octree, angle, blur, desc+epilogue (s)
};

double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void compute_pyramid(Extractor &E, const uint8_t *img, int w, int h, int stride) {
    E.lv.assign(E.p.nlevels, Level());
    for (int l = 0; l < E.p.nlevels; l++) {
        Level &L = E.lv[l];
        float scale = E.p.inv_sf[l];
        L.w = cv_round_f((float)w * scale); L.h = cv_round_f((float)h * scale);
        L.stride = L.w + 2 * kEdge;
        L.buf.assign((size_t)L.stride * (L.h + 2 * kEdge), 0);
        uint8_t *in = L.buf.data() + (size_t)kEdge * L.stride + kEdge;
        if (l == 0) for (int y = 0; y < h; y++) memcpy(in + (size_t)y * L.stride, img + (size_t)y * stride, w);
        else {
            Level &P = E.lv[l - 1];
            resize_linear(P.buf.data() + (size_t)kEdge * P.stride + kEdge, P.w, P.h, P.stride, in, L.w, L.h, L.stride);
        }
        fill_border_reflect101(L.buf.data(), L.w, L.h, L.stride, kEdge);
    }
}

void detect_level(Extractor &E, int l) {
    Level &L = E.lv[l];
    const uint8_t *in = L.buf.data() + (size_t)kEdge * L.stride + kEdge;
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    const float W = 35;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    L.cand.clear();
    if (nCols <= 0 || nRows <= 0) return;   // reference would divide by zero; defined as "no cells"
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    std::vector<FastPt> cell;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
            const uint8_t *roi = in + (size_t)y0 * L.stride + x0;
            fast9_16(roi, cw, chh, L.stride, E.p.iniTh, cell);
            if (cell.empty()) fast9_16(roi, cw, chh, L.stride, E.p.minTh, cell);
            for (auto &k : cell) L.cand.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
        }
    }
}

void select_level(Extractor &E, int l) {
    Level &L = E.lv[l];
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    std::vector<int> sel;
    L.kps.clear();
    if (L.cand.empty()) return;   // reference: DistributeOctTree on an empty set returns nothing
    distribute_octtree(L.cand, minBX, maxBX, minBY, maxBY, E.p.nfeat[l], sel);
    const int scaledPatch = (int)(kPatchSize * E.p.sf[l]);
    for (int k : sel) {
        KeyPt kp;
        kp.x = (float)L.cand[k].x + minBX; kp.y = (float)L.cand[k].y + minBY;
        kp.size = (float)scaledPatch; kp.angle = -1; kp.response = (float)L.cand[k].score;
        kp.octave = l; kp.class_id = -1;
        L.kps.push_back(kp);
    }
}

int extract(Extractor &E, const uint8_t *img, int w, int h, int stride, int lap0, int lap1, int fma_mode) {
    E.out_kps.clear(); E.out_desc.clear(); E.n_mono = 0;
    if (!img || w <= 0 || h <= 0) return -1;   // :994-995
    double t0 = now_s();
    compute_pyramid(E, img, w, h, stride);
    double t1 = now_s();
    for (int l = 0; l < E.p.nlevels; l++) detect_level(E, l);
    double t2 = now_s();
    for (int l = 0; l < E.p.nlevels; l++) select_level(E, l);
    double t3 = now_s();
    for (int l = 0; l < E.p.nlevels; l++) {
        Level &L = E.lv[l];
        const uint8_t *in = L.buf.data() + (size_t)kEdge * L.stride + kEdge;
        for (auto &kp : L.kps) kp.angle = ic_angle(in, L.stride, cv_round_f(kp.x), cv_round_f(kp.y), E.p.umax);
    }
    double t4 = now_s();
    int n = 0;
    for (auto &L : E.lv) n += (int)L.kps.size();
    E.out_kps.assign(n, KeyPt());
    E.out_desc.assign((size_t)n * 32, 0);
    int mono = 0, stereo = n - 1;
    double tb = 0;
    for (int l = 0; l < E.p.nlevels; l++) {
        Level &L = E.lv[l];
        if (L.kps.empty()) continue;
        double tb0 = now_s();
        L.blur.assign((size_t)L.w * L.h, 0);
        blur7(L.buf.data() + (size_t)kEdge * L.stride + kEdge, L.w, L.h, L.stride, L.blur.data(), L.w);
        L.blurred = true;
        tb += now_s() - tb0;
        float scale = E.p.sf[l];
        for (auto &kp0 : L.kps) {
            uint8_t d[32];
            orb_descriptor(L.blur.data(), L.w, cv_round_f(kp0.x), cv_round_f(kp0.y), kp0.angle, fma_mode, d);
            KeyPt kp = kp0;
            if (l != 0) { kp.x *= scale; kp.y *= scale; }
            int slot;
            if (kp.x >= lap0 && kp.x <= lap1) slot = stereo--; else slot = mono++;
            E.out_kps[slot] = kp;
            memcpy(&E.out_desc[(size_t)slot * 32], d, 32);
        }
    }
    double t5 = now_s();
    E.n_mono = mono;
    E.t_stage[0] += t1 - t0; E.t_stage[1] += t2 - t1; E.t_stage[2] += t3 - t2;
    E.t_stage[3] += t4 - t3; E.t_stage[4] += tb; E.t_stage[5] += (t5 - t4) - tb;
    return mono;
}

// ORBmatcher::DescriptorDistance.  ORBmatcher.cc:2577-2591 (bit-twiddling popcount over 8 x u32).
int descriptor_distance(const uint8_t *a, const uint8_t *b) {
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32); memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// Frame grid.  Frame.h:24-25, Frame.cc:541-582 (AssignFeaturesToGrid), 969-978 (PosInGrid),
// 890-967 (GetFeaturesInArea).  One grid per camera, camera-local indices.
const int kGridCols = 64, kGridRows = 48;
struct Grid {
    float minX, minY, invW, invH;
    std::vector<int> cell[kGridCols][kGridRows];
};
void grid_build(Grid &g, const KeyPt *kps, int n, float minX, float maxX, float minY, float maxY) {
    g.minX = minX; g.minY = minY;
    g.invW = (float)kGridCols / (maxX - minX);
    g.invH = (float)kGridRows / (maxY - minY);
    for (int i = 0; i < kGridCols; i++) for (int j = 0; j < kGridRows; j++) g.cell[i][j].clear();
    for (int i = 0; i < n; i++) {
        int px = (int)roundf((kps[i].x - minX) * g.invW);
        int py = (int)roundf((kps[i].y - minY) * g.invH);
        if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
        g.cell[px][py].push_back(i);
    }
}
void features_in_area(const Grid &g, const KeyPt *kps, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int> &out) {
    out.clear();
    const int c0 = std::max(0, (int)floorf((x - g.minX - r) * g.invW));
    if (c0 >= kGridCols) return;
    const int c1 = std::min(kGridCols - 1, (int)ceilf((x - g.minX + r) * g.invW));
    if (c1 < 0) return;
    const int r0 = std::max(0, (int)floorf((y - g.minY - r) * g.invH));
    if (r0 >= kGridRows) return;
    const int r1 = std::min(kGridRows - 1, (int)ceilf((y - g.minY + r) * g.invH));
    if (r1 < 0) return;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = c0; ix <= c1; ix++)
        for (int iy = r0; iy <= r1; iy++)
            for (int idx : g.cell[ix][iy]) {
                const KeyPt &kp = kps[idx];
                if (checkLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                if (fabsf(kp.x - x) < r && fabsf(kp.y - y) < r) out.push_back(idx);
            }
}

}  // namespace

// ================================================================================================
// C interface (ctypes).  Everything is prefixed orc_ ("oracle").
// ================================================================================================
extern "C" {

struct orc_kp { float x, y, size, angle, response; int octave, class_id; };

void *orc_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    Extractor *E = new Extractor();
    make_params(E->p, nfeatures, scaleFactor, nlevels, iniTh, minTh);
    return E;
}
void orc_destroy(void *h) { delete (Extractor *)h; }

void orc_tables(void *h, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2, int *nfeat, int *umax) {
    Extractor *E = (Extractor *)h;
    for (int i = 0; i < E->p.nlevels; i++) {
        sf[i] = E->p.sf[i]; inv_sf[i] = E->p.inv_sf[i]; sigma2[i] = E->p.sigma2[i];
        inv_sigma2[i] = E->p.inv_sigma2[i]; nfeat[i] = E->p.nfeat[i];
    }
    for (int v = 0; v <= kHalfPatch; v++) umax[v] = E->p.umax[v];
}

int orc_extract(void *h, const uint8_t *img, int w, int hh, int stride, int lap0, int lap1, int fma_mode) {
    return extract(*(Extractor *)h, img, w, hh, stride, lap0, lap1, fma_mode);
}
int orc_num_keypoints(void *h) { return (int)((Extractor *)h)->out_kps.size(); }
void orc_result(void *h, orc_kp *kps, uint8_t *desc) {
    Extractor *E = (Extractor *)h;
    memcpy(kps, E->out_kps.data(), E->out_kps.size() * sizeof(KeyPt));
    memcpy(desc, E->out_desc.data(), E->out_desc.size());
}
void orc_level_size(void *h, int l, int *w, int *hh, int *stride) {
    Level &L = ((Extractor *)h)->lv[l]; *w = L.w; *hh = L.h; *stride = L.stride;
}
void orc_level_bordered(void *h, int l, uint8_t *dst) {   // (h+38) x (w+38), tightly packed
    Level &L = ((Extractor *)h)->lv[l]; memcpy(dst, L.buf.data(), L.buf.size());
}
int orc_level_blur(void *h, int l, uint8_t *dst) {        // h x w tightly packed; 0 if level not blurred
    Level &L = ((Extractor *)h)->lv[l]; if (!L.blurred) return 0; memcpy(dst, L.blur.data(), L.blur.size()); return 1;
}
int orc_level_num_candidates(void *h, int l) { return (int)((Extractor *)h)->lv[l].cand.size(); }
void orc_level_candidates(void *h, int l, int *xs, int *ys, int *scores) {
    Level &L = ((Extractor *)h)->lv[l];
    for (size_t i = 0; i < L.cand.size(); i++) { xs[i] = L.cand[i].x; ys[i] = L.cand[i].y; scores[i] = L.cand[i].score; }
}
int orc_level_num_keypoints(void *h, int l) { return (int)((Extractor *)h)->lv[l].kps.size(); }
void orc_level_keypoints(void *h, int l, orc_kp *kps) {
    Level &L = ((Extractor *)h)->lv[l]; memcpy(kps, L.kps.data(), L.kps.size() * sizeof(KeyPt));
}
void orc_stage_seconds(void *h, double *t6, int reset) {
    Extractor *E = (Extractor *)h;
    for (int i = 0; i < 6; i++) { t6[i] = E->t_stage[i]; if (reset) E->t_stage[i] = 0; }
}

// --- primitives, individually testable against cv2 -------------------------------------------------
void orc_resize_linear(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride) {
    resize_linear(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_border_reflect101(uint8_t *buf, int w, int h, int stride, int b) { fill_border_reflect101(buf, w, h, stride, b); }
int orc_fast(const uint8_t *img, int w, int h, int stride, int threshold, int *xs, int *ys, int *scores, int cap) {
    std::vector<FastPt> out; fast9_16(img, w, h, stride, threshold, out);
    int n = (int)out.size();
    for (int i = 0; i < n && i < cap; i++) { xs[i] = out[i].x; ys[i] = out[i].y; scores[i] = out[i].score; }
    return n;
}
void orc_blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) { blur7(src, w, h, sstride, dst, dstride); }
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_fast_atan2_n(const float *y, const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]); }
float orc_ic_angle(const uint8_t *img, int stride, int x, int y) {
    Params p; make_params(p, 1000, 1.2f, 8, 20, 7);
    return ic_angle(img, stride, x, y, p.umax);
}
void orc_descriptor(const uint8_t *img, int stride, int x, int y, float angle, int fma_mode, uint8_t *desc) {
    orb_descriptor(img, stride, x, y, angle, fma_mode, desc);
}
// candidate selection alone: pts (x,y,score) border-relative -> selected indices in output order
int orc_distribute(const int *xs, const int *ys, const int *scores, int n, int minX, int maxX, int minY, int maxY,
                   int N, int *sel, int cap) {
    std::vector<FastPt> pts(n);
    for (int i = 0; i < n; i++) pts[i] = {xs[i], ys[i], scores[i]};
    std::vector<int> r;
    if (n > 0) distribute_octtree(pts, minX, maxX, minY, maxY, N, r);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) sel[i] = r[i];
    return (int)r.size();
}
// std::sort with compareNodes on (size, UL.x) pairs: returns the permutation (for the introsort emulation test)
void orc_sort_nodes(const int *sizes, const int *ulx, int n, int *perm) {
    std::vector<Node> nodes(n);
    std::vector<std::pair<int, Node *>> v(n);
    for (int i = 0; i < n; i++) { nodes[i].ulx = ulx[i]; v[i] = std::make_pair(sizes[i], &nodes[i]); }
    std::sort(v.begin(), v.end(), compare_nodes);
    for (int i = 0; i < n; i++) perm[i] = (int)(v[i].second - nodes.data());
}

// glibc cosf/sinf as the reference calls them (ORBextractor.cc:48) — for the device restatement test
void orc_libm_sincosf_n(const float *x, float *s, float *c, int n) { for (int i = 0; i < n; i++) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); } }

// --- batch extraction on host threads (CPU baseline: one std::thread per image, Frame.cc:1841-1862) ---
// imgs: n images, each w x h tightly packed.  Returns total keypoints; per-image counts in n_kp/n_mono.
int orc_extract_batch(int nimg, const uint8_t *imgs, int w, int h, int nfeatures, float sf, int nlevels, int iniTh,
                      int minTh, int lap0, int lap1, int fma_mode, int nthreads, int *n_kp, int *n_mono) {
    std::atomic<int> next(0), total(0);
    auto work = [&]() {
        Extractor E; make_params(E.p, nfeatures, sf, nlevels, iniTh, minTh);
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= nimg) break;
            int m = extract(E, imgs + (size_t)i * w * h, w, h, w, lap0, lap1, fma_mode);
            if (n_kp) n_kp[i] = (int)E.out_kps.size();
            if (n_mono) n_mono[i] = m;
            total += (int)E.out_kps.size();
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, nthreads); t++) th.emplace_back(work);
    for (auto &t : th) t.join();
    return total.load();
}

// --- matching -----------------------------------------------------------------------------------------
int orc_hamming(const uint8_t *a, const uint8_t *b) { return descriptor_distance(a, b); }

// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) (Frame.cc:1427,1483): per query the two nearest train rows,
// ordered (distance asc, trainIdx asc).  idx/dist are [nq][2]; missing entries are -1.
void orc_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt, int *idx, int *dist, int nthreads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i0 = next.fetch_add(64);
            if (i0 >= nq) break;
            for (int i = i0; i < std::min(nq, i0 + 64); i++) {
                int b0 = 1 << 30, b1 = 1 << 30, i0x = -1, i1x = -1;
                for (int j = 0; j < nt; j++) {
                    int d = descriptor_distance(q + (size_t)i * 32, t + (size_t)j * 32);
                    if (d < b0) { b1 = b0; i1x = i0x; b0 = d; i0x = j; }
                    else if (d < b1) { b1 = d; i1x = j; }
                }
                idx[2 * i] = i0x; dist[2 * i] = i0x < 0 ? -1 : b0;
                idx[2 * i + 1] = i1x; dist[2 * i + 1] = i1x < 0 ? -1 : b1;
            }
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < std::max(1, nthreads); k++) th.emplace_back(work);
    for (auto &x : th) x.join();
}

// Lowe ratio test as the callers write it (Frame.cc:1434 ratio 0.7, :1491 ratio 0.8): float distances,
// the product evaluated in double.  keep[i] = 1 iff two neighbours exist and d0 < d1 * ratio.
void orc_ratio_test(const int *dist, int nq, double ratio, uint8_t *keep) {
    for (int i = 0; i < nq; i++)
        keep[i] = (dist[2 * i] >= 0 && dist[2 * i + 1] >= 0 && (float)dist[2 * i] < (float)dist[2 * i + 1] * ratio) ? 1 : 0;
}

// Frame grid (per camera)
void *orc_grid_create(const orc_kp *kps, int n, float minX, float maxX, float minY, float maxY) {
    Grid *g = new Grid(); grid_build(*g, (const KeyPt *)kps, n, minX, maxX, minY, maxY); return g;
}
void orc_grid_destroy(void *g) { delete (Grid *)g; }
int orc_grid_query(void *g, const orc_kp *kps, float x, float y, float r, int minLevel, int maxLevel, int *out, int cap) {
    std::vector<int> v; features_in_area(*(Grid *)g, (const KeyPt *)kps, x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints), ORBmatcher.cc:23-340,
// over plain arrays.  Frame side: ncam cameras (1 = the Nleft==-1 monocular/rectified layout, 2 or 4 = the
// multi-camera layout with Nleft..Nsideright), camera c owns global rows [cam_off[c], cam_off[c+1]).
//   kps / desc            : all keypoints (camera-local order inside each block) and their descriptor rows
//   grids[c]              : orc_grid_create() of camera c's keypoints
//   sf                    : F.mvScaleFactors
//   occupied[idx]         : 1 iff F.mvpMapPoints[idx] is set with Observations()>0 on entry (global idx)
//   l2r / r2l             : mvLeftToRightMatch (size Nleft) / mvRightToLeftMatch (size Nright), or NULL
//   u_right               : F.mvuRight (only read when ncam==1), or NULL
// Map-point side (nmp points; per-camera arrays are [nmp][4]):
//   in_view, proj_x, proj_y, level, view_cos ; proj_xr[nmp] (rectified right u) ; depth[nmp]; bad[nmp];
//   has_obs[nmp] (Observations()>0 — what makes a claim visible to later points); mp_desc [nmp][32]
// Output: assigned[idx] = map-point index written to F.mvpMapPoints[idx] (last writer), -1 if untouched.
// Returns nmatches.
int orc_search_by_projection(int ncam, const int *cam_off, const orc_kp *kps, const uint8_t *desc, void **grids,
                             const float *sf, const uint8_t *occupied, const int *l2r, const int *r2l,
                             const float *u_right, int nmp, const uint8_t *in_view, const float *proj_x,
                             const float *proj_y, const int *level, const float *view_cos, const float *proj_xr,
                             const float *depth, const uint8_t *bad, const uint8_t *has_obs, const uint8_t *mp_desc,
                             float th, int bFarPoints, float thFarPoints, float nnratio, int *assigned) {
    const int N = cam_off[ncam];
    std::vector<uint8_t> occ(occupied, occupied + N);
    for (int i = 0; i < N; i++) assigned[i] = -1;
    const int TH_HIGH = 100;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> cand;
    for (int m = 0; m < nmp; m++) {
        const uint8_t *iv = in_view + 4 * m;
        if (!iv[0] && !iv[1] && !iv[2] && !iv[3]) continue;
        if (bFarPoints && depth[m] > thFarPoints) continue;
        if (bad[m]) continue;
        bool skip_rest = false;
        for (int c = 0; c < 4 && !skip_rest; c++) {
            if (!iv[c]) continue;
            if (c >= 1 && ncam < 2) continue;           // F.Nleft == -1
            if (c >= 2 && ncam < 4) continue;           // F.Nsideleft == -1
            const int lvl = level[4 * m + c];
            if (c >= 1 && lvl == -1) continue;
            float r = view_cos[4 * m + c] > 0.998 ? 2.5f : 4.0f;   // RadiusByViewingCos :342-347
            if (c == 0 && bFactor) r *= th;                        // only camera 0 applies th (:51-52)
            const KeyPt *ck = (const KeyPt *)kps + cam_off[c];
            features_in_area(*(Grid *)grids[c], ck, proj_x[4 * m + c], proj_y[4 * m + c], r * sf[lvl], lvl - 1, lvl, cand);
            if (cand.empty()) continue;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (int idx : cand) {
                const int g = idx + cam_off[c];
                if (occ[g]) continue;
                if (c == 0 && ncam == 1 && u_right && u_right[idx] > 0) {
                    const float er = fabsf(proj_xr[m] - u_right[idx]);
                    if (er > r * sf[lvl]) continue;
                }
                const int dist = descriptor_distance(mp_desc + (size_t)m * 32, desc + (size_t)g * 32);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
                    bestLevel = ck[idx].octave; bestIdx = idx;
                } else if (dist < bestDist2) { bestLevel2 = ck[idx].octave; bestDist2 = dist; }
            }
            if (bestDist <= TH_HIGH) {
                if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) { skip_rest = true; continue; }  // `continue` of the MapPoint loop
                const int g = bestIdx + cam_off[c];
                if (c == 0) {
                    assigned[g] = m; occ[g] = has_obs[m];
                    if (ncam >= 2 && l2r && l2r[bestIdx] != -1) {
                        const int gr = l2r[bestIdx] + cam_off[1];
                        assigned[gr] = m; occ[gr] = has_obs[m];
                        nmatches++;
                    }
                    nmatches++;
                } else if (c == 1) {
                    if (r2l && r2l[bestIdx] != -1) {
                        const int gl = r2l[bestIdx];
                        assigned[gl] = m; occ[gl] = has_obs[m];
                        nmatches++;
                    }
                    assigned[g] = m; occ[g] = has_obs[m];
                    nmatches++;
                } else {
                    assigned[g] = m; occ[g] = has_obs[m];
                    nmatches++;
                }
            }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), ORBmatcher.cc:1985-2413, over plain
// arrays (frame side as in orc_search_by_projection).  Per last-frame keypoint i: valid[i] (pMP && !outlier && invzc >= 0 && camera-0
// projection in bounds, :2008-2026), proj_x / proj_y [npts][4] (what the reference hands to GetFeaturesInArea for camera c: the side
// cameras get the LEFT uv, :2250-2271), proj_ur[i] = uv.x - mbf * invzc, octave / angle of the last keypoint, has_obs, descriptor.
// Output: assigned[idx] = i written to CurrentFrame.mvpMapPoints[idx] (-1 = untouched or cleared by the rotation filter).
int orc_search_by_projection_last(int ncam, const int *cam_off, const orc_kp *kps, const uint8_t *desc, void **grids, const float *sf,
                                  const uint8_t *occupied, const float *u_right, int npts, const uint8_t *valid, const float *proj_x,
                                  const float *proj_y, const float *proj_ur, const int *last_octave, const float *last_angle,
                                  const uint8_t *has_obs, const uint8_t *pt_desc, float th, int levelGate,
                                  int checkOrientation, int maxDist, int *assigned) {
    // levelGate: 1 = bForward, 2 = bBackward, 0 = neither; 3 = [l - 1, l], the gate of the KeyFrame / Sim3 overloads (:668-893)
    const int bForward = levelGate == 1, bBackward = levelGate == 2;
    // maxDist = TH_HIGH (:2103); the relocalisation overload (Frame, KeyFrame, sAlreadyFound, th, ORBdist) (:2415-2535) has the same
    // body for one camera with ORBdist in its place (:2487)
    const int N = cam_off[ncam];
    const int HISTO_LENGTH = 30;
    std::vector<uint8_t> occ(N, 0);
    if (occupied) occ.assign(occupied, occupied + N);
    for (int i = 0; i < N; i++) assigned[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> cand;
    for (int i = 0; i < npts; i++) {
        if (!valid[i]) continue;
        const int nLastOctave = last_octave[i];
        const float radius = th * sf[nLastOctave];                                              // :2040
        for (int c = 0; c < ncam; c++) {
            const KeyPt *ck = (const KeyPt *)kps + cam_off[c];
            const float u = proj_x[4 * i + c], v = proj_y[4 * i + c];
            if (bForward) features_in_area(*(Grid *)grids[c], ck, u, v, radius, nLastOctave, -1, cand);
            else if (bBackward) features_in_area(*(Grid *)grids[c], ck, u, v, radius, 0, nLastOctave, cand);
            else if (levelGate == 3) {   // KeyFrame::GetFeaturesInArea has no level arguments; the loop gates kpLevel itself (:759); ncam == 1
                std::vector<int> all;
                features_in_area(*(Grid *)grids[c], ck, u, v, radius, -1, -1, all);
                cand.clear();
                for (int i2 : all) if (!(ck[i2].octave < nLastOctave - 1 || ck[i2].octave > nLastOctave)) cand.push_back(i2);
            }
            else features_in_area(*(Grid *)grids[c], ck, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
            if (c == 0 && cand.empty()) break;                                                  // :2062 `continue` of the point loop
            int bestDist = 256, bestIdx2 = -1;
            for (int i2 : cand) {
                const int g = i2 + cam_off[c];
                if (occ[g]) continue;
                if (c == 0 && ncam == 1 && u_right && u_right[i2] > 0) {                          // :2079-2084
                    const float er = fabsf(proj_ur[i] - u_right[i2]);
                    if (er > radius) continue;
                }
                const int dist = descriptor_distance(pt_desc + (size_t)i * 32, desc + (size_t)g * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
            }
            if (bestDist <= maxDist) {
                const int g = bestIdx2 + cam_off[c];
                assigned[g] = i; occ[g] = has_obs ? has_obs[i] : 1;
                nmatches++;
                if (checkOrientation) {
                    float rot = last_angle[i] - ck[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(g);
                }
            }
        }
    }
    if (checkOrientation) {                                                                     // :2392-2410 + ComputeThreeMaxima :2537-2573
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int g : rotHist[i]) { assigned[g] = -1; nmatches--; }
    }
    return nmatches;
}

// ---- window search without claims: the search loop of Fuse (ORBmatcher.cc:1549-1601, 1719-1738) and SearchBySim3 (:1837-1861, 1911-1935)
// grid / kps: one camera of the KeyFrame; KeyFrame::GetFeaturesInArea (KeyFrame.cc:771-834) = the Frame query without level gate.
void orc_window_best(void *grid, const orc_kp *kps, const uint8_t *desc, const uint8_t *skip, const float *u_right, const float *invSigma2,
                     int nq, const uint8_t *valid, const float *qx, const float *qy, const float *q_ur, const float *radius,
                     const int *min_level, const int *max_level, const uint8_t *q_desc, int *best_idx, int *best_dist) {
    const KeyPt *ck = (const KeyPt *)kps;
    std::vector<int> cand;
    for (int q = 0; q < nq; q++) {
        int bestDist = 256, bestIdx = -1;
        if (valid[q]) {
            features_in_area(*(Grid *)grid, ck, qx[q], qy[q], radius[q], -1, -1, cand);
            for (int idx : cand) {
                if (skip && skip[idx]) continue;
                const int kpLevel = ck[idx].octave;
                if (kpLevel < min_level[q] || kpLevel > max_level[q]) continue;
                if (invSigma2) {
                    if (u_right && u_right[idx] >= 0) {                                  // :1571-1583
                        const float ex = qx[q] - ck[idx].x, ey = qy[q] - ck[idx].y, er = q_ur[q] - u_right[idx];
                        const float e2 = ex * ex + ey * ey + er * er;
                        if (e2 * invSigma2[kpLevel] > 7.8) continue;
                    } else {                                                             // :1584-1592
                        const float ex = qx[q] - ck[idx].x, ey = qy[q] - ck[idx].y;
                        const float e2 = ex * ex + ey * ey;
                        if (e2 * invSigma2[kpLevel] > 5.99) continue;
                    }
                }
                const int dist = descriptor_distance(q_desc + (size_t)q * 32, desc + (size_t)idx * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
        }
        best_idx[q] = bestIdx; best_dist[q] = bestDist;
    }
}

// ---- ORBmatcher::SearchBySim3, ORBmatcher.cc:1771-1983: the two searches and the agreement check over plain arrays
int orc_search_by_sim3(void *grid1, const orc_kp *kps1, const uint8_t *desc1, int n1, void *grid2, const orc_kp *kps2, const uint8_t *desc2,
                       int n2, const uint8_t *valid1, const float *u1, const float *v1, const float *radius1, const int *level1,
                       const uint8_t *mpdesc1, const uint8_t *valid2, const float *u2, const float *v2, const float *radius2,
                       const int *level2, const uint8_t *mpdesc2, int *match12) {
    const int TH_HIGH = 100;
    std::vector<int> vnMatch1(n1, -1), vnMatch2(n2, -1), cand;
    const KeyPt *k1 = (const KeyPt *)kps1, *k2 = (const KeyPt *)kps2;
    for (int i1 = 0; i1 < n1; i1++) {
        if (!valid1[i1]) continue;
        features_in_area(*(Grid *)grid2, k2, u1[i1], v1[i1], radius1[i1], -1, -1, cand);
        int bestDist = INT_MAX, bestIdx = -1;
        for (int idx : cand) {
            if (k2[idx].octave < level1[i1] - 1 || k2[idx].octave > level1[i1]) continue;
            const int dist = descriptor_distance(mpdesc1 + (size_t)i1 * 32, desc2 + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_HIGH) vnMatch1[i1] = bestIdx;
    }
    for (int i2 = 0; i2 < n2; i2++) {
        if (!valid2[i2]) continue;
        features_in_area(*(Grid *)grid1, k1, u2[i2], v2[i2], radius2[i2], -1, -1, cand);
        int bestDist = INT_MAX, bestIdx = -1;
        for (int idx : cand) {
            if (k1[idx].octave < level2[i2] - 1 || k1[idx].octave > level2[i2]) continue;
            const int dist = descriptor_distance(mpdesc2 + (size_t)i2 * 32, desc1 + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_HIGH) vnMatch2[i2] = bestIdx;
    }
    int nFound = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    return nFound;
}

// ORBmatcher::SearchForTriangulation Hamming core, ORBmatcher.cc:1218-1427, with the vocabulary FeatureVectors
// given as CSR (sorted node ids + index lists) and the geometry test as an optional pair mask.
//   nodes1[nn1], ptr1[nn1+1], idx1[...] : KF1 feature vector ; same for KF2
//   skip1[idx] / skip2[idx] : keypoint already has a MapPoint (or fails bOnlyStereo) -> skipped
//   pair_ok : NULL (bCoarse=true) or a callback standing for the epipole-distance + epipolarConstrain tests
// Output: match12[n1] = idx2 or -1.  Returns nmatches.  (vbMatched2 is never set by the reference, :1204/1261.)
typedef int (*orc_pair_pred)(int idx1, int idx2, void *user);
int orc_search_for_triangulation(int nn1, const int *nodes1, const int *ptr1, const int *idx1v, int nn2,
                                 const int *nodes2, const int *ptr2, const int *idx2v, const uint8_t *desc1, int n1,
                                 const uint8_t *desc2, const uint8_t *skip1, const uint8_t *skip2,
                                 orc_pair_pred pair_ok, void *user, int *match12) {
    const int TH_LOW = 50;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = ptr1[a]; i1 < ptr1[a + 1]; i1++) {
                const int id1 = idx1v[i1];
                if (skip1[id1]) continue;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = ptr2[b]; i2 < ptr2[b + 1]; i2++) {
                    const int id2 = idx2v[i2];
                    if (skip2[id2]) continue;
                    const int dist = descriptor_distance(desc1 + (size_t)id1 * 32, desc2 + (size_t)id2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (pair_ok && !pair_ok(id1, id2, user)) continue;
                    bestIdx2 = id2; bestDist = dist;
                }
                if (bestIdx2 >= 0) { match12[id1] = bestIdx2; nmatches++; }
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) a++;   // lower_bound
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) b++;
        }
    }
    return nmatches;
}


// ---- ORBmatcher::SearchForTriangulation with every gate of the candidate loop, ORBmatcher.cc:1131-1456 (the variant LocalMapping
// calls: bCoarse = false, src/LocalMapping.cc:465-468) ---------------------------------------------------------------------------------
// ncam = 1 (NLeft == -1: monocular / rectified stereo, no mpCamera2), 2 (left + right) or 4 (+ side-left, side-right); camera c of
// keyframe k owns the feature indices [cam_offk[c], cam_offk[c+1]).  has_mp = GetMapPoint(idx) != NULL; u_right = mvuRight (only read
// when ncam == 1; NULL = all -1); ep = epipole of KF1's centre in KF2's first camera (:1141-1147); sf2 = pKF2->mvScaleFactors.
// The reference carries (pCamera1, pCamera2, R12, t12) ACROSS candidates: a camera pair without a branch in :1301-1378 keeps what the
// previous candidate that got that far had set.  `state` numbers the pairs: 0 = the single pair of the ncam == 1 case (T12, :1156-1160),
// 1 ll, 2 lr, 3 rl, 4 rr (:1301-1330), 5 l-sl, 6 sl-l, 7 sl-sl, 8 r-sr, 9 sr-r, 10 sr-sr (:1332-1377).  With ncam >= 2 the reference
// starts with pCamera1/2 = the two left cameras and R12 / t12 UNINITIALISED (:1152-1154): here the initial state is 1 (ll), which is
// what the reference computes whenever the first candidate that reaches the epipolar test has a branch of its own.
// pred(idx1, idx2, state, user) stands for pCamera1->epipolarConstrain(pCamera2, kp1, kp2, R12, t12, sigma1, sigma2) (:1380-1387).
// match12[n1] = idx2 or -1.  Returns nmatches.
typedef int (*orc_pair_state_pred)(int idx1, int idx2, int state, void *user);
static int three_maxima_filter(std::vector<int> *rotHist, int *match);
static int rot_bin(float a1, float a2);
int orc_search_for_triangulation_full(int ncam, const int *cam_off1, const orc_kp *kps1, const uint8_t *desc1, const uint8_t *has_mp1,
                                      const float *u_right1, int nn1, const int *nodes1, const int *ptr1, const int *idx1v,
                                      const int *cam_off2, const orc_kp *kps2, const uint8_t *desc2, const uint8_t *has_mp2,
                                      const float *u_right2, int nn2, const int *nodes2, const int *ptr2, const int *idx2v, const float *ep,
                                      const float *sf2, int bOnlyStereo, int bCoarse, int checkOrientation, orc_pair_state_pred pred, void *user,
                                      int *match12) {
    const int TH_LOW = 50;
    const KeyPt *k1 = (const KeyPt *)kps1, *k2 = (const KeyPt *)kps2;
    const int n1 = cam_off1[ncam];
    for (int i = 0; i < n1; i++) match12[i] = -1;
    auto cam_of = [&](const int *off, int idx) { int c = 0; while (c + 1 < ncam && idx >= off[c + 1]) c++; return c; };
    std::vector<int> rotHist[30];
    int nmatches = 0, a = 0, b = 0;
    int state = ncam == 1 ? 0 : 1;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = ptr1[a]; i1 < ptr1[a + 1]; i1++) {
                const int id1 = idx1v[i1];
                if (has_mp1[id1]) continue;
                const bool bStereo1 = ncam == 1 && u_right1 && u_right1[id1] >= 0;     // (!pKF1->mpCamera2 && mvuRight >= 0) :1228
                if (bOnlyStereo && !bStereo1) continue;
                const int cam1 = cam_of(cam_off1, id1);
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = ptr2[b]; i2 < ptr2[b + 1]; i2++) {
                    const int id2 = idx2v[i2];
                    if (has_mp2[id2]) continue;                                          // vbMatched2 is never set (:1204, 1261)
                    const bool bStereo2 = ncam == 1 && u_right2 && u_right2[id2] >= 0;
                    if (bOnlyStereo && !bStereo2) continue;
                    const int dist = descriptor_distance(desc1 + (size_t)id1 * 32, desc2 + (size_t)id2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const int cam2 = cam_of(cam_off2, id2);
                    if (!bStereo1 && !bStereo2 && ncam == 1) {                            // :1292-1299
                        const float distex = ep[0] - k2[id2].x, distey = ep[1] - k2[id2].y;
                        if (distex * distex + distey * distey < 100 * sf2[k2[id2].octave]) continue;
                    }
                    if (ncam >= 2) {                                                      // :1301-1378
                        if (cam1 == 1 && cam2 == 1) state = 4;
                        else if (cam1 == 1 && cam2 == 0) state = 3;
                        else if (cam1 == 0 && cam2 == 1) state = 2;
                        else if (cam1 == 0 && cam2 == 0) state = 1;
                        if (ncam >= 4) {
                            if (cam1 == 0 && cam2 == 2) state = 5;
                            else if (cam1 == 2 && cam2 == 0) state = 6;
                            else if (cam1 == 2 && cam2 == 2) state = 7;
                            else if (cam1 == 1 && cam2 == 3) state = 8;
                            else if (cam1 == 3 && cam2 == 1) state = 9;
                            else if (cam1 == 3 && cam2 == 3) state = 10;
                        }
                    }
                    if (bCoarse || pred(id1, id2, state, user)) { bestIdx2 = id2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[id1] = bestIdx2; nmatches++;
                    if (checkOrientation) rotHist[rot_bin(k1[id1].angle, k2[bestIdx2].angle)].push_back(id1);
                }
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) a++;   // lower_bound
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) b++;
        }
    }
    if (checkOrientation) nmatches -= three_maxima_filter(rotHist, match12);
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), ORBmatcher.cc:349-666, over plain arrays ------------------------
// FeatureVectors as CSR (sorted node ids, index lists); skipKF[i] = !pMP || pMP->isBad(); angles by feature index; ncamF = 1 for
// F.Nleft == -1, else 2 / 4 with camOffF = {0, Nleft, Nleft+Nright, +Nsideleft, +Nsideright}.  matchF[nF] = realIdxKF or -1.
static int three_maxima_filter(std::vector<int> *rotHist, int *match) {   // :637-663 + ComputeThreeMaxima :2537-2573
    const int L = 30;
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1, removed = 0;
    for (int i = 0; i < L; i++) {
        const int sz = (int)rotHist[i].size();
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
        else if (sz > max3) { max3 = sz; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
    for (int i = 0; i < L; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (int j : rotHist[i]) { match[j] = -1; removed++; }
    }
    return removed;
}
static int rot_bin(float a1, float a2) {
    const float factor = 1.0f / 30;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == 30) bin = 0;
    return bin;
}

// ORBmatcher::ComputeThreeMaxima (:2537-2573) on a histogram given by its bin sizes
void orc_three_maxima(const int *sizes, int L, int *ind3) {
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3v = -1;
    for (int i = 0; i < L; i++) {
        const int sz = sizes[i];
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3v = ind2; ind2 = ind1; ind1 = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; ind3v = ind2; ind2 = i; }
        else if (sz > max3) { max3 = sz; ind3v = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3v = -1; }
    else if (max3 < 0.1f * (float)max1) ind3v = -1;
    ind3[0] = ind1; ind3[1] = ind2; ind3[2] = ind3v;
}

int orc_search_by_bow_frame(int nnKF, const int *nodesKF, const int *ptrKF, const int *idxKF, int nnF, const int *nodesF, const int *ptrF,
                            const int *idxF, const uint8_t *descKF, const uint8_t *descF, int nF, const uint8_t *skipKF, const float *angleKF,
                            const float *angleF, int ncamF, const int *camOffF, float nnratio, int checkOrientation, int *matchF) {
    const int TH_LOW = 50;
    for (int i = 0; i < nF; i++) matchF[i] = -1;
    std::vector<int> rotHist[30];
    int nmatches = 0, a = 0, b = 0;
    while (a < nnKF && b < nnF) {
        if (nodesKF[a] == nodesF[b]) {
            for (int iKF = ptrKF[a]; iKF < ptrKF[a + 1]; iKF++) {
                const int realIdxKF = idxKF[iKF];
                if (skipKF && skipKF[realIdxKF]) continue;
                const uint8_t *dKF = descKF + (size_t)realIdxKF * 32;
                int bestDist1[4] = {256, 256, 256, 256}, bestIdx[4] = {-1, -1, -1, -1}, bestDist2[4] = {256, 256, 256, 256};
                for (int iF = ptrF[b]; iF < ptrF[b + 1]; iF++) {
                    const int realIdxF = idxF[iF];
                    if (matchF[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dKF, descF + (size_t)realIdxF * 32);
                    int c = 0;
                    if (ncamF > 1) {
                        c = -1;
                        for (int k = 0; k < ncamF; k++) if (realIdxF >= camOffF[k] && realIdxF < camOffF[k + 1]) c = k;
                        if (c < 0) continue;
                    }
                    if (dist < bestDist1[c]) { bestDist2[c] = bestDist1[c]; bestDist1[c] = dist; bestIdx[c] = realIdxF; }
                    else if (dist < bestDist2[c]) bestDist2[c] = dist;
                }
                if (bestDist1[0] <= TH_LOW) {
                    for (int c = 0; c < (ncamF > 1 ? ncamF : 1); c++) {
                        if (c == 0) { if (!((float)bestDist1[0] < nnratio * (float)bestDist2[0])) continue; }
                        else if (!(bestDist1[c] <= TH_LOW)) continue;                       // ratio test disabled by `|| true`
                        matchF[bestIdx[c]] = realIdxKF;
                        if (checkOrientation) rotHist[rot_bin(angleKF[realIdxKF], angleF[bestIdx[c]])].push_back(bestIdx[c]);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (nodesKF[a] < nodesF[b]) {
            while (a < nnKF && nodesKF[a] < nodesF[b]) a++;                                 // lower_bound
        } else {
            while (b < nnF && nodesF[b] < nodesKF[a]) b++;
        }
    }
    if (checkOrientation) nmatches -= three_maxima_filter(rotHist, matchF);
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), ORBmatcher.cc:1006-1129 -------------------------------------------
int orc_search_by_bow_keyframes(int nn1, const int *nodes1, const int *ptr1, const int *idx1v, int nn2, const int *nodes2, const int *ptr2,
                                const int *idx2v, const uint8_t *desc1, int n1, const uint8_t *desc2, int n2, const uint8_t *skip1,
                                const uint8_t *skip2, const float *angle1, const float *angle2, float nnratio, int checkOrientation,
                                int *match12) {
    const int TH_LOW = 50;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<uint8_t> vbMatched2(n2, 0);
    std::vector<int> rotHist[30];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = ptr1[a]; i1 < ptr1[a + 1]; i1++) {
                const int idx1 = idx1v[i1];
                if (skip1 && skip1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = ptr2[b]; i2 < ptr2[b + 1]; i2++) {
                    const int idx2 = idx2v[i2];
                    if (vbMatched2[idx2] || (skip2 && skip2[idx2])) continue;
                    const int dist = descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        match12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = 1;
                        if (checkOrientation) rotHist[rot_bin(angle1[idx1], angle2[bestIdx2])].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) a++;
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) b++;
        }
    }
    if (checkOrientation) nmatches -= three_maxima_filter(rotHist, match12);
    return nmatches;
}

// ---- ORBmatcher::SearchForInitialization, ORBmatcher.cc:895-1004, over plain arrays (grid2 = orc_grid_create of F2.mvKeysUn) -------
int orc_search_for_initialization(const orc_kp *kps1, const uint8_t *desc1, int n1, void *grid2, const orc_kp *kps2, const uint8_t *desc2,
                                  int n2, float *vbPrevMatched, int windowSize, float nnratio, int checkOrientation, int *vnMatches12) {
    const int TH_LOW = 50;
    const KeyPt *k1 = (const KeyPt *)kps1, *k2 = (const KeyPt *)kps2;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[30];
    std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1), vIndices2;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = k1[i1].octave;
        if (level1 > 0) continue;
        features_in_area(*(Grid *)grid2, k2, vbPrevMatched[2 * i1], vbPrevMatched[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOrientation) rotHist[rot_bin(k1[i1].angle, k2[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOrientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; i++) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < 30; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) { vbPrevMatched[2 * i1] = k2[vnMatches12[i1]].x; vbPrevMatched[2 * i1 + 1] = k2[vnMatches12[i1]].y; }
    return nmatches;
}

// ---- Frame::ComputeStereoMatches (src/Frame.cc:1057-1219) -------------------------------------------------------
// hL / hR: oracle extractors that hold the last extraction of the (rectified) left / right image: mvKeys = out_kps of hL,
// mvKeysRight = out_kps of hR, mDescriptors(Right) = out_desc, mvImagePyramid = lv[].  mb = baseline, mbf = baseline * fx.
// Outputs mvuRight[N], mvDepth[N] (-1 = no match).  Returns the number of keypoints with depth.
int orc_compute_stereo_matches(void *hL, void *hR, float mb, float mbf, float *mvuRight, float *mvDepth) {
    Extractor &EL = *(Extractor *)hL, &ER = *(Extractor *)hR;
    const int N = (int)EL.out_kps.size();
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }             // :1058-1059
    const int thOrbDist = (100 + 50) / 2;                                                 // :1061
    const int nRows = EL.lv[0].h;                                                         // :1063
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    const int Nr = (int)ER.out_kps.size();
    for (int iR = 0; iR < Nr; iR++) {                                                     // :1073-1083
        const KeyPt &kp = ER.out_kps[iR];
        const float kpY = kp.y;
        const float r = 2.0f * ER.p.sf[kp.octave];
        const int maxr = (int)ceil(kpY + r);
        const int minr = (int)floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);                     // (the reference indexes unchecked)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;                                   // :1086-1088
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {                                                      // :1094-1201
        const KeyPt &kpL = EL.out_kps[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const std::vector<size_t> &vCandidates = vRowIndices[(int)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;                                                               // TH_HIGH
        size_t bestIdxR = 0;
        const uint8_t *dL = EL.out_desc.data() + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const KeyPt &kpR = ER.out_kps[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, ER.out_desc.data() + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                                                       // :1138
            const float uR0 = ER.out_kps[bestIdxR].x;
            const float scaleFactor = EL.p.inv_sf[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const Level &PL = EL.lv[kpL.octave], &PR = ER.lv[kpL.octave];
            const uint8_t *IL = PL.buf.data() + (size_t)(kEdge + (int)(scaledvL - w)) * PL.stride + kEdge + (int)(scaleduL - w);
            int bestDistS = INT32_MAX, bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const uint8_t *IR = PR.buf.data() + (size_t)(kEdge + (int)(scaledvL - w)) * PR.stride + kEdge + (int)(scaleduR0 + incR - w);
                int sad = 0;                                                              // cv::norm(IL, IR, NORM_L1) on 8-bit mats
                for (int y = 0; y < 2 * w + 1; y++)
                    for (int x = 0; x < 2 * w + 1; x++) sad += abs((int)IL[(size_t)y * PL.stride + x] - (int)IR[(size_t)y * PR.stride + x]);
                const float dist = (float)(double)sad;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = EL.p.sf[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;                                                       // (the reference reads vDistIdx[0] unchecked)
    std::sort(vDistIdx.begin(), vDistIdx.end());                                          // :1203-1205
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {                                 // :1207-1214
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1; mvDepth[vDistIdx[i].second] = -1; kept--;
    }
    return kept;
}


// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483), batched over map points ---------------------
// Group g holds the observed descriptors rows [ptr[g], ptr[g+1]) of `desc` (in the order the reference pushes them into
// vDescriptors).  best[g] = index inside the group of the descriptor with the least median distance to the others
// (first minimum; median = sorted row [0.5 * (N - 1)]), or -1 for an empty group.
void orc_distinctive_descriptors(const uint8_t *desc, const int *ptr, int ngroups, int *best) {
    for (int g = 0; g < ngroups; g++) {
        const int N = ptr[g + 1] - ptr[g];
        best[g] = -1;
        if (N <= 0) continue;
        const uint8_t *d = desc + (size_t)ptr[g] * 32;
        std::vector<float> Distances((size_t)N * N);                                    // :452-461
        for (int i = 0; i < N; i++) {
            Distances[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int distij = descriptor_distance(d + (size_t)i * 32, d + (size_t)j * 32);
                Distances[(size_t)i * N + j] = (float)distij; Distances[(size_t)j * N + i] = (float)distij;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;                                           // :464-475
        for (int i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + (size_t)i * N, Distances.begin() + (size_t)(i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best[g] = BestIdx;
    }
}


// ---- DBoW2 TemplatedVocabulary<FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259) ----------
// Vocabulary tree as flat arrays: node i has children child[child_ptr[i] .. child_ptr[i+1]) (node ids, in the order of
// Node::children), a 32-byte descriptor, and, for leaves (no children), word_id / weight.  Node 0 is the root.
// Per feature (:1213-1259): descend taking the FIRST child of minimum FORB::distance (FORB.cpp:81-101 == popcount);
// nid = node reached at level m_L - levelsup (0 if that level is <= 0 or never reached).
void orc_vocab_transform(const int *child_ptr, const int *child, const uint8_t *node_desc, const int *word_id, int L, const uint8_t *feat,
                         int n, int levelsup, int *out_word, int *out_node) {
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t *d = feat + (size_t)f * 32;
        int nid = 0, final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int b = child_ptr[final_id], e = child_ptr[final_id + 1];
            final_id = child[b];
            double best_d = descriptor_distance(d, node_desc + (size_t)final_id * 32);
            for (int k = b + 1; k < e; k++) {
                const int id = child[k];
                const double dd = descriptor_distance(d, node_desc + (size_t)id * 32);
                if (dd < best_d) { best_d = dd; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_ptr[final_id + 1] > child_ptr[final_id]);
        out_word[f] = word_id[final_id];
        out_node[f] = nid;
    }
}

// BowVector / FeatureVector of the batch transform (:1127-1193) for TF_IDF weighting + L1 scoring (the ORBvoc settings):
// v.addWeight(id, w) in feature order, then v.normalize(L1) (BowVector.cpp:34-88); fv.addFeature(nid, i) (FeatureVector.cpp:31-45).
// Outputs are the maps flattened in key order: bow_ids / bow_vals [return value], fv_nodes / fv_ptr / fv_idx.
int orc_bow_from_words(const int *word, const int *node, const double *weight, int n, int *bow_ids, double *bow_vals, int *fv_nodes,
                       int *fv_ptr, int *fv_idx, int *n_fv_nodes) {
    std::vector<std::pair<int, double>> v;                     // sorted by id, like the std::map
    std::vector<std::pair<int, std::vector<unsigned>>> fv;
    for (int i = 0; i < n; i++) {
        const double w = weight[i];
        if (!(w > 0)) continue;                                // stopped word
        auto it = std::lower_bound(v.begin(), v.end(), word[i], [](const std::pair<int, double> &a, int id) { return a.first < id; });
        if (it != v.end() && it->first == word[i]) it->second += w; else v.insert(it, std::make_pair(word[i], w));
        auto jt = std::lower_bound(fv.begin(), fv.end(), node[i], [](const std::pair<int, std::vector<unsigned>> &a, int id) { return a.first < id; });
        if (jt == fv.end() || jt->first != node[i]) jt = fv.insert(jt, std::make_pair(node[i], std::vector<unsigned>()));
        jt->second.push_back((unsigned)i);
    }
    double norm = 0.0;
    for (auto &e : v) norm += fabs(e.second);
    if (norm > 0.0) for (auto &e : v) e.second /= norm;
    for (size_t k = 0; k < v.size(); k++) { bow_ids[k] = v[k].first; bow_vals[k] = v[k].second; }
    int p = 0;
    for (size_t k = 0; k < fv.size(); k++) {
        fv_nodes[k] = fv[k].first; fv_ptr[k] = p;
        for (unsigned x : fv[k].second) fv_idx[p++] = (int)x;
    }
    fv_ptr[fv.size()] = p;
    *n_fv_nodes = (int)fv.size();
    return (int)v.size();
}

// ---- Fisheye stereo post-filter: KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:319-388) and its
// caller Frame::ComputeStereoFishEyeMatches / ComputeMultiFishEyeMatches (src/Frame.cc:1409-1459, 1461-1525) ---------------------
// fp32 as written in the reference (this file is built with -ffp-contract=off).  The one third-party piece is
// Eigen::JacobiSVD<Matrix4f>(A, ComputeFullV).matrixV().col(3) (:422-424; Eigen is not in this image): restated as the
// eigenvector of the smallest eigenvalue of A^T A by cyclic Jacobi in fp64 (pinned against numpy.linalg.svd in
// tests/test_oracle_fisheye.py), rounded to fp32.
struct OrcKb8Stereo { float cam1[8], cam2[8], R12[9], t12[3], precision1, precision2; };

static void kb8_unproject(const float *P, float precision, float u, float v, float ray[3]) {   // :96-126
    const float pwx = (u - P[2]) / P[0], pwy = (v - P[3]) / P[1];
    float scale = 1.f;
    float theta_d = sqrtf(pwx * pwx + pwy * pwy);
    theta_d = fminf(fmaxf((float)(-M_PI / 2.f), theta_d), (float)(M_PI / 2.f));
    if (theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
            float k0_theta2 = P[4] * theta2, k1_theta4 = P[5] * theta4;
            float k2_theta6 = P[6] * theta6, k3_theta8 = P[7] * theta8;
            float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                              (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < precision) break;
        }
        scale = tanf(theta) / theta_d;
    }
    ray[0] = pwx * scale; ray[1] = pwy * scale; ray[2] = 1.f;
}

static void kb8_project(const float *P, const float X[3], float uv[2]) {   // :48-65
    const float x2_plus_y2 = X[0] * X[0] + X[1] * X[1];
    const float theta = atan2f(sqrtf(x2_plus_y2), X[2]);
    const float psi = atan2f(X[1], X[0]);
    const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2, theta9 = theta7 * theta2;
    const float r = theta + P[4] * theta3 + P[5] * theta5 + P[6] * theta7 + P[7] * theta9;
    uv[0] = P[0] * r * cosf(psi) + P[2];
    uv[1] = P[1] * r * sinf(psi) + P[3];
}

// v4 = unit right singular vector of the smallest singular value of the row-major 4x4 matrix a
static void null_vector4(const float *a, double v4[4]) {
    double M[4][4], V[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += (double)a[4 * k + i] * (double)a[4 * k + j];
            M[i][j] = s; V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) off += M[p][q] * M[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                if (M[p][q] == 0.0) continue;
                const double th = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; k++) { const double kp = M[k][p], kq = M[k][q]; M[k][p] = c * kp - sn * kq; M[k][q] = sn * kp + c * kq; }
                for (int k = 0; k < 4; k++) { const double pk = M[p][k], qk = M[q][k]; M[p][k] = c * pk - sn * qk; M[q][k] = sn * pk + c * qk; }
                for (int k = 0; k < 4; k++) { const double kp = V[k][p], kq = V[k][q]; V[k][p] = c * kp - sn * kq; V[k][q] = sn * kp + c * kq; }
            }
    }
    int best = 0;
    for (int j = 1; j < 4; j++) if (M[j][j] < M[best][best]) best = j;
    for (int i = 0; i < 4; i++) v4[i] = V[i][best];
}

// returns z1 or the reject code; margin (if not NULL) = smallest relative distance of a tested quantity from its gate
static float kb8_triangulate_matches(const OrcKb8Stereo &S, float u1, float v1, float u2, float v2, float sigmaLevel, float unc,
                                     float p3D[3], float *margin) {
    float mg = 1e30f;
    auto gate = [&](double value, double limit) { const float m = (float)(fabs(value - limit) / (fabs(limit) + 1e-12)); if (m < mg) mg = m; };
    auto done = [&](float code) { if (margin) *margin = mg; return code; };
    float r1[3], r2[3];
    kb8_unproject(S.cam1, S.precision1, u1, v1, r1);
    kb8_unproject(S.cam2, S.precision2, u2, v2, r2);
    const float *R = S.R12;
    float r21[3];
    for (int i = 0; i < 3; i++) r21[i] = R[3 * i] * r2[0] + R[3 * i + 1] * r2[1] + R[3 * i + 2] * r2[2];
    const float n1 = sqrtf(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]), n21 = sqrtf(r21[0] * r21[0] + r21[1] * r21[1] + r21[2] * r21[2]);
    const float cosParallaxRays = (r1[0] * r21[0] + r1[1] * r21[1] + r1[2] * r21[2]) / (n1 * n21);
    gate(1.0 - (double)cosParallaxRays, 1.0 - 0.9998);
    if (cosParallaxRays > 0.9998) return done(-1);
    float Tcw2[3][4];
    for (int i = 0; i < 3; i++) {
        Tcw2[i][0] = R[i]; Tcw2[i][1] = R[3 + i]; Tcw2[i][2] = R[6 + i];
        Tcw2[i][3] = -R[i] * S.t12[0] + -R[3 + i] * S.t12[1] + -R[6 + i] * S.t12[2];
    }
    float A[16] = {-1.f, 0.f, r1[0], 0.f, 0.f, -1.f, r1[1], 0.f};                       // p1 * Tcw1.row(2) - Tcw1.row(0 / 1), Tcw1 = [I|0]
    for (int j = 0; j < 4; j++) {
        A[8 + j] = r2[0] * Tcw2[2][j] - Tcw2[0][j];
        A[12 + j] = r2[1] * Tcw2[2][j] - Tcw2[1][j];
    }
    double v4[4];
    null_vector4(A, v4);
    const float hw = (float)v4[3];
    float x3D[3] = {(float)v4[0] / hw, (float)v4[1] / hw, (float)v4[2] / hw};
    const float z1 = x3D[2];
    if (z1 <= 0) return done(-2);
    const float z2 = Tcw2[2][0] * x3D[0] + Tcw2[2][1] * x3D[1] + Tcw2[2][2] * x3D[2] + Tcw2[2][3];
    if (z2 <= 0) return done(-3);
    float uv1[2];
    kb8_project(S.cam1, x3D, uv1);
    const float errX1 = uv1[0] - u1, errY1 = uv1[1] - v1;
    gate(errX1 * errX1 + errY1 * errY1, 5.991 * sigmaLevel);
    if ((errX1 * errX1 + errY1 * errY1) > 5.991 * sigmaLevel) return done(-4);
    float x3D2[3] = {Tcw2[0][0] * x3D[0] + Tcw2[0][1] * x3D[1] + Tcw2[0][2] * x3D[2] + Tcw2[0][3],
                     Tcw2[1][0] * x3D[0] + Tcw2[1][1] * x3D[1] + Tcw2[1][2] * x3D[2] + Tcw2[1][3], z2};
    float uv2[2];
    kb8_project(S.cam2, x3D2, uv2);
    const float errX2 = uv2[0] - u2, errY2 = uv2[1] - v2;
    gate(errX2 * errX2 + errY2 * errY2, 5.991 * unc);
    if ((errX2 * errX2 + errY2 * errY2) > 5.991 * unc) return done(-5);
    gate(z1, 0.0001);
    p3D[0] = x3D[0]; p3D[1] = x3D[1]; p3D[2] = x3D[2];
    return done(z1);
}

void orc_null_vector4(const float *a16, double *v4) { null_vector4(a16, v4); }

void orc_kb8_unproject(const float *cam8, float precision, const float *xy, int n, float *rays) {
    for (int i = 0; i < n; i++) kb8_unproject(cam8, precision, xy[2 * i], xy[2 * i + 1], rays + 3 * i);
}
void orc_kb8_project(const float *cam8, const float *xyz, int n, float *uv) {
    for (int i = 0; i < n; i++) kb8_project(cam8, xyz + 3 * i, uv + 2 * i);
}

void orc_kb8_triangulate_matches(const OrcKb8Stereo *rig, const float *xy1, const float *xy2, const float *sigma1, const float *sigma2,
                                 int n, float *depth, float *p3d, float *margin) {
    for (int i = 0; i < n; i++) {
        float p[3] = {0.f, 0.f, 0.f};
        depth[i] = kb8_triangulate_matches(*rig, xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1], sigma1[i], sigma2[i], p,
                                           margin ? margin + i : nullptr);
        p3d[3 * i] = p[0]; p3d[3 * i + 1] = p[1]; p3d[3 * i + 2] = p[2];
    }
}

// Frame::ComputeStereoFishEyeMatches (:1409-1459).  margin[n_left]: gate margin of the query's triangulation (1e30 where the
// ratio test already failed).  Returns nMatches; *desc_matches = descMatches.
int orc_stereo_fisheye_matches(const OrcKb8Stereo *rig, const orc_kp *kpsL, const uint8_t *descL, int nL, int monoL, const orc_kp *kpsR,
                               const uint8_t *descR, int nR, int monoR, const float *levelSigma2, double ratio, int *l2r, int *r2l,
                               float *depth, float *p3d, float *margin, int *desc_matches) {
    for (int i = 0; i < nL; i++) { l2r[i] = -1; depth[i] = -1.0f; p3d[3 * i] = p3d[3 * i + 1] = p3d[3 * i + 2] = 0.f; if (margin) margin[i] = 1e30f; }
    for (int i = 0; i < nR; i++) r2l[i] = -1;
    int nMatches = 0, descMatches = 0;
    const int nq = nL - monoL, nt = nR - monoR;
    std::vector<int> idx((size_t)nq * 2 + 2), dist((size_t)nq * 2 + 2);
    if (nq > 0) orc_knn2(descL + (size_t)monoL * 32, nq, descR + (size_t)monoR * 32, nt, idx.data(), dist.data(), 1);
    for (int q = 0; q < nq; q++) {
        if (idx[2 * q] < 0 || idx[2 * q + 1] < 0) continue;                                         // (*it).size() >= 2
        if (!((float)dist[2 * q] < (float)dist[2 * q + 1] * ratio)) continue;                        // :1434 (float * double)
        descMatches++;
        const int iL = q + monoL, iR = idx[2 * q] + monoR;
        const float sigma1 = levelSigma2[kpsL[iL].octave], sigma2 = levelSigma2[kpsR[iR].octave];
        float p[3];
        const float d = kb8_triangulate_matches(*rig, kpsL[iL].x, kpsL[iL].y, kpsR[iR].x, kpsR[iR].y, sigma1, sigma2, p, margin ? margin + iL : nullptr);
        if (d > 0.0001f) {
            l2r[iL] = iR; r2l[iR] = iL;
            p3d[3 * iL] = p[0]; p3d[3 * iL + 1] = p[1]; p3d[3 * iL + 2] = p[2];
            depth[iL] = d;
            nMatches++;
        }
    }
    if (desc_matches) *desc_matches = descMatches;
    return nMatches;
}


// ---- SearchForTriangulation with the reference's two epipolar predicates built in (test counterpart of
// orbx_search_for_triangulation_full): pairs[11] = {model (0 pinhole, 1 Kannala-Brandt), reserved, F12[9], OrcKb8Stereo} per carried
// state of orc_search_for_triangulation_full.
struct OrcEpipolarPair { int model, reserved; float F12[9]; OrcKb8Stereo kb8; };
// Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:103-132) given F12 = K1^-T [t12]x R12 K2^-1 (:112-115)
static bool pinhole_epipolar(const float *F, float x1, float y1, float x2, float y2, float unc) {
    const float a = x1 * F[0] + y1 * F[3] + F[6];
    const float b = x1 * F[1] + y1 * F[4] + F[7];
    const float c = x1 * F[2] + y1 * F[5] + F[8];
    const float num = a * x2 + b * y2 + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc;
}
int orc_pinhole_epipolar(const float *F, float x1, float y1, float x2, float y2, float unc) { return pinhole_epipolar(F, x1, y1, x2, y2, unc) ? 1 : 0; }
struct SftCamsCtx { const OrcEpipolarPair *pairs; const KeyPt *k1, *k2; const float *s1, *s2; std::vector<float> *margins; };
static int sft_cams_pred(int i1, int i2, int state, void *user) {
    SftCamsCtx &C = *(SftCamsCtx *)user;
    const OrcEpipolarPair &P = C.pairs[state];
    const KeyPt &a = C.k1[i1], &b = C.k2[i2];
    if (P.model == 1) {     // KannalaBrandt8::epipolarConstrain (src/CameraModels/KannalaBrandt8.cpp:219-230)
        float p3[3]; float margin = 0;
        const float z = kb8_triangulate_matches(P.kb8, a.x, a.y, b.x, b.y, C.s1[a.octave], C.s2[b.octave], p3, &margin);
        if (C.margins) C.margins->push_back(margin);
        return z > 0.0001f;
    }
    return pinhole_epipolar(P.F12, a.x, a.y, b.x, b.y, C.s2[b.octave]);
}
int orc_search_for_triangulation_cams(int ncam, const int *cam_off1, const orc_kp *kps1, const uint8_t *desc1, const uint8_t *has_mp1,
                                      const float *u_right1, int nn1, const int *nodes1, const int *ptr1, const int *idx1v,
                                      const int *cam_off2, const orc_kp *kps2, const uint8_t *desc2, const uint8_t *has_mp2,
                                      const float *u_right2, int nn2, const int *nodes2, const int *ptr2, const int *idx2v, const float *ep,
                                      const float *sf2, const float *sigma2_1, const float *sigma2_2, int bOnlyStereo, int bCoarse,
                                      int checkOrientation, const OrcEpipolarPair *pairs, int *match12, float *min_margin) {
    std::vector<float> margins;
    SftCamsCtx C{pairs, (const KeyPt *)kps1, (const KeyPt *)kps2, sigma2_1, sigma2_2, min_margin ? &margins : nullptr};
    const int n = orc_search_for_triangulation_full(ncam, cam_off1, kps1, desc1, has_mp1, u_right1, nn1, nodes1, ptr1, idx1v, cam_off2, kps2, desc2,
                                                    has_mp2, u_right2, nn2, nodes2, ptr2, idx2v, ep, sf2, bOnlyStereo, bCoarse, checkOrientation,
                                                    sft_cams_pred, &C, match12);
    if (min_margin) { float mm = 1e30f; for (float v : margins) mm = std::min(mm, v); *min_margin = mm; }
    return n;
}

}  // extern "C"
