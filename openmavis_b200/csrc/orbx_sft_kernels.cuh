// orbx_sft_kernels.cuh — ORBmatcher::SearchForTriangulation with the geometric tests inside the candidate loop
// (src/ORBmatcher.cc:1131-1456 with bCoarse = false, the variant LocalMapping::CreateNewMapPoints calls, src/LocalMapping.cc:465-468).
//
// Row = one unmatched idx1 of a vocabulary node shared by the two keyframes; its candidates are the node's idx2 list, visited in
// order.  A candidate that passes `dist <= min(TH_LOW, best so far)` (:1274) and the epipole-distance reject of the monocular case
// (:1292-1299) selects the camera pair (pCamera1, pCamera2, R12, t12) of its two cameras (:1301-1378) and must pass
// pCamera1->epipolarConstrain (:1380-1387) to become the row's best: the LAST such minimum wins.
// The pair selection has no branch for some camera combinations (left / side-right, right / side-left, ...): such a candidate is tested
// with whatever pair the previous candidate that got that far had selected — state carried across candidates, rows and nodes.
//   k_sft_rows : one thread per row walks its candidates exactly like the reference, with the carried state UNKNOWN on entry; a row
//                whose walk never needs the unknown state is final and reports the state it leaves behind (or "unchanged");
//   k_sft_fix  : only if some row needed the carried state: one thread walks the rows in order, threads the state through the rows'
//                exit states and re-walks just the unresolved rows with their true entry state;
//   k_sft_rot  : the rotation-histogram filter (:1405-1446, only with mbCheckOrientation).
// Predicates: Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:103-132) on a caller-supplied F12 (the caller evaluates
// K1^-T [t12]x R12 K2^-1 once per camera pair with the reference's own expression), fp32 with separately rounded operations — bit-exact;
// KannalaBrandt8::epipolarConstrain = TriangulateMatches(...) > 0.0001 (src/CameraModels/KannalaBrandt8.cpp:219-230) through
// kb8_triangulate_matches of orbx_fisheye_kernels.cuh (float geometry with atan2f / tanf: tolerance-based parity).
#pragma once

#include "orbx_common.cuh"
#include "orbx_fisheye_kernels.cuh"

namespace orbx {

constexpr int kSftStates = 11;      // 0: the ncam == 1 pair; 1 ll, 2 lr, 3 rl, 4 rr, 5 l-sl, 6 sl-l, 7 sl-sl, 8 r-sr, 9 sr-r, 10 sr-sr
constexpr int kSftUnknown = -1;

struct SftArgs {
    const uint4 *d1, *d2;            // descriptors
    const orbx_kp *k1, *k2;
    int nrows;
    const int *row_q, *row_begin, *row_len, *idx2;
    const uint8_t *skip2;            // has a MapPoint, or fails bOnlyStereo
    const uint8_t *st1, *st2;        // bStereo1 / bStereo2 (:1228, 1265): only the ncam == 1 layout can set them
    int ncam;
    int off1[5], off2[5];
    float ep[2];
    float sf2[16], sigma1[16], sigma2[16];
    int coarse;
    const orbx_epipolar_pair *pairs; // [kSftStates], device
    int *match12;                    // [n1], pre-filled with -1
    int *exit_state;                 // [nrows]
    int *unresolved;                 // [nrows + 1]: per-row flag, [nrows] = any
};

__device__ __forceinline__ int sft_cam_of(const int *off, int ncam, int idx) {
    int c = 0;
    while (c + 1 < ncam && idx >= off[c + 1]) c++;
    return c;
}

// camera pair selected by (camera of idx1, camera of idx2), or kSftUnknown when the reference has no branch for it (:1301-1378)
__device__ __forceinline__ int sft_pair_state(int ncam, int c1, int c2) {
    if (ncam < 2) return 0;
    if (c1 <= 1 && c2 <= 1) return c1 == 0 ? (c2 == 0 ? 1 : 2) : (c2 == 0 ? 3 : 4);
    if (ncam >= 4) {
        if (c1 == 0 && c2 == 2) return 5;
        if (c1 == 2 && c2 == 0) return 6;
        if (c1 == 2 && c2 == 2) return 7;
        if (c1 == 1 && c2 == 3) return 8;
        if (c1 == 3 && c2 == 1) return 9;
        if (c1 == 3 && c2 == 3) return 10;
    }
    return kSftUnknown;
}

// Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:118-132) given F12
__device__ __forceinline__ bool pinhole_epipolar(const float *F, float x1, float y1, float x2, float y2, float unc) {
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, x2), __fmul_rn(b, y2)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.0f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < 3.84 * (double)unc;
}

__device__ inline bool sft_epipolar(const orbx_epipolar_pair &P, const orbx_kp &a, const orbx_kp &b, float s1, float s2) {
    if (P.model == ORBX_CAM_KB8) {
        float p3[3];
        return kb8_triangulate_matches(P.kb8, a.x, a.y, b.x, b.y, s1, s2, p3) > 0.0001f;
    }
    return pinhole_epipolar(P.F12, a.x, a.y, b.x, b.y, s2);
}

// Walks one row.  `state` = the carried pair on entry (kSftUnknown in the parallel pass).  Returns false if the walk needed an unknown
// carried state (nothing is written then); otherwise best = idx2 or -1 and `state` = the pair left behind.
__device__ inline bool sft_walk_row(const SftArgs &A, int r, int &state, int &best_out) {
    const int i1 = A.row_q[r];
    const uint4 qa = A.d1[2 * i1], qb = A.d1[2 * i1 + 1];
    const orbx_kp kp1 = A.k1[i1];
    const int c1 = sft_cam_of(A.off1, A.ncam, i1);
    const bool stereo1 = A.st1[i1] != 0;
    int best = 50 /* TH_LOW */, bi = -1;
    const int b = A.row_begin[r], e = b + A.row_len[r];
    for (int k = b; k < e; k++) {
        const int c = A.idx2[k];
        if (A.skip2[c]) continue;
        const int d = hamming256(qa, qb, A.d2[2 * c], A.d2[2 * c + 1]);
        if (d > 50 || d > best) continue;
        const orbx_kp kp2 = A.k2[c];
        if (!stereo1 && !A.st2[c] && A.ncam == 1) {      // epipole-distance reject (:1292-1299)
            const float ex = __fsub_rn(A.ep[0], kp2.x), ey = __fsub_rn(A.ep[1], kp2.y);
            if (__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < __fmul_rn(100.0f, A.sf2[kp2.octave])) continue;
        }
        const int s = sft_pair_state(A.ncam, c1, sft_cam_of(A.off2, A.ncam, c));
        if (s != kSftUnknown) state = s;
        else if (state == kSftUnknown) return false;
        if (A.coarse || sft_epipolar(A.pairs[state], kp1, kp2, A.sigma1[kp1.octave], A.sigma2[kp2.octave])) { bi = c; best = d; }
    }
    best_out = bi;
    return true;
}

__global__ void __launch_bounds__(128) k_sft_rows(const __grid_constant__ SftArgs A) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.nrows) return;
    int state = A.ncam == 1 ? 0 : kSftUnknown, best = -1;
    if (sft_walk_row(A, r, state, best)) {
        A.match12[A.row_q[r]] = best;
        A.exit_state[r] = state;
        A.unresolved[r] = 0;
    } else {
        A.unresolved[r] = 1;
        A.unresolved[A.nrows] = 1;
    }
}

// ordered fix-up, one thread (launched only when a row stayed unresolved)
__global__ void k_sft_fix(const __grid_constant__ SftArgs A) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int state = 1;     // pCamera1 / pCamera2 start as the two left cameras (:1154); R12 / t12 are uninitialised there in the reference
    for (int r = 0; r < A.nrows; r++) {
        if (!A.unresolved[r]) {
            if (A.exit_state[r] != kSftUnknown) state = A.exit_state[r];
            continue;
        }
        int best = -1;
        sft_walk_row(A, r, state, best);
        A.match12[A.row_q[r]] = best;
    }
}

// rotation histogram filter (:1405-1446): bins of round((angle1 - angle2) / 30), keep the three largest (ComputeThreeMaxima :2537-2573)
__global__ void __launch_bounds__(256) k_sft_rot(const orbx_kp *__restrict__ k1, const orbx_kp *__restrict__ k2, int n1, int *__restrict__ match12) {
    __shared__ int hist[30], keep[3];
    if (threadIdx.x < 30) hist[threadIdx.x] = 0;
    __syncthreads();
    auto bin_of = [&](int i, int j) {
        float rot = __fsub_rn(k1[i].angle, k2[j].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
        return bin == 30 ? 0 : bin;
    };
    for (int i = threadIdx.x; i < n1; i += blockDim.x)
        if (match12[i] >= 0) atomicAdd(&hist[bin_of(i, match12[i])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n1; i += blockDim.x)
        if (match12[i] >= 0) {
            const int b = bin_of(i, match12[i]);
            if (b != keep[0] && b != keep[1] && b != keep[2]) match12[i] = -1;
        }
}

}  // namespace orbx
