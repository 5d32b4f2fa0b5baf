// orbx_stereo_kernels.cuh — Frame::ComputeStereoMatches (src/Frame.cc:1057-1219) on device-resident extraction results:
// row-band Hamming matching of left against right keypoints of a rectified pair, 11x11 SAD refinement on the pyramid level of
// the left keypoint (mvImagePyramid of both extractors: the planes the extraction just produced), parabola sub-pixel fit,
// depth, and the median-based outlier removal.  SURVEY.md §8(f) N2.
#pragma once

#include "orbx_extract_kernels.cuh"

namespace orbx {

struct StereoArgs {
    const orbx_kp *kpL, *kpR;        // final keypoints of the left / right image (mvKeys, mvKeysRight)
    const uint4 *descL, *descR;
    const int *nL, *nR;              // keypoint counts (device)
    const uint8_t *pyrL, *pyrR;      // pyramid arenas of the extractor(s) that hold the two images (same geometry)
    int imgL, imgR;                  // plane indices of the two images
    float inv_sf[kMaxLevels];        // mvInvScaleFactors
    float maxD, mbf;                 // mbf / mb, mbf
    float *uRight, *depth;           // [cap] outputs
    int *sad;                        // [cap] best SAD of the accepted matches, -1 otherwise
};

// one warp per left keypoint
__global__ void __launch_bounds__(256) k_stereo_match(Geom G, StereoArgs A) {
    const int lane = threadIdx.x & 31;
    const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int N = *A.nL, Nr = *A.nR;
    if (iL >= N) return;
    if (lane == 0) { A.uRight[iL] = -1.0f; A.depth[iL] = -1.0f; A.sad[iL] = -1; }
    const orbx_kp kl = A.kpL[iL];
    const int levelL = kl.octave;
    const float vL = kl.y, uL = kl.x;
    const int rowL = (int)vL;                                   // vRowIndices[vL]
    const float minU = __fsub_rn(uL, A.maxD), maxU = uL;        // minD = 0
    if (maxU < 0) return;
    const uint4 qa = A.descL[2 * iL], qb = A.descL[2 * iL + 1];
    // ---- best right keypoint of the row band: first minimum in ascending iR order, strictly below TH_HIGH (:1117-1135)
    uint32_t best = (100u << 16) | 0xFFFFu;
    for (int iR = lane; iR < Nr; iR += 32) {
        const orbx_kp kr = A.kpR[iR];
        if (kr.octave < levelL - 1 || kr.octave > levelL + 1) continue;
        const float r = __fmul_rn(2.0f, G.lv[kr.octave].scale);
        const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));   // :1076-1080
        if (rowL < minr || rowL > maxr) continue;
        if (!(kr.x >= minU && kr.x <= maxU)) continue;
        const uint32_t d = (uint32_t)hamming256(qa, qb, A.descR[2 * iR], A.descR[2 * iR + 1]);
        best = min(best, (d << 16) | (uint32_t)iR);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    const int bestDist = (int)(best >> 16);
    if (bestDist >= 100 || bestDist >= (100 + 50) / 2) return;  // no candidate below TH_HIGH, or not below thOrbDist (:1138)
    const int bestIdxR = (int)(best & 0xFFFFu);
    // ---- sub-pixel match by correlation (:1139-1199)
    const float uR0 = A.kpR[bestIdxR].x;
    const float scaleFactor = A.inv_sf[levelL];
    const float scaleduL = roundf(__fmul_rn(kl.x, scaleFactor)), scaledvL = roundf(__fmul_rn(kl.y, scaleFactor));
    const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
    const int w = 5, Lr = 5;
    const LevelGeom &P = G.lv[levelL];
    const float iniu = scaleduR0 + Lr - w, endu = scaleduR0 + Lr + w + 1;
    if (iniu < 0 || endu >= (float)P.w) return;
    const uint8_t *IL = plane_ptr(A.pyrL, P, A.imgL) + (long long)(kYO + (int)(scaledvL - w)) * P.pitch + kXO + (int)(scaleduL - w);
    const uint8_t *IR = plane_ptr(A.pyrR, P, A.imgR) + (long long)(kYO + (int)(scaledvL - w)) * P.pitch + kXO + (int)(scaleduR0 - Lr - w);
    // lane = (row y = lane / 3, column group): 11 rows x 3 lanes; each lane sums its columns for the 11 shifts
    int sadv[11];
#pragma unroll
    for (int s = 0; s < 11; s++) sadv[s] = 0;
    if (lane < 11) {   // one lane per patch row keeps the code simple: 11 left pixels, 21 right pixels
        int l[11], r[21];
#pragma unroll
        for (int x = 0; x < 11; x++) l[x] = IL[(long long)lane * P.pitch + x];
#pragma unroll
        for (int x = 0; x < 21; x++) r[x] = IR[(long long)lane * P.pitch + x];
#pragma unroll
        for (int s = 0; s < 11; s++)
#pragma unroll
            for (int x = 0; x < 11; x++) sadv[s] += abs(l[x] - r[x + s]);
    }
#pragma unroll
    for (int s = 0; s < 11; s++)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sadv[s] += __shfl_xor_sync(0xFFFFFFFFu, sadv[s], o);   // lanes 0..15 hold the rows
    if (lane != 0) return;
    int bestS = INT_MAX, bestincR = 0;
#pragma unroll
    for (int s = 0; s < 11; s++)
        if ((float)sadv[s] < (float)bestS) { bestS = sadv[s]; bestincR = s - Lr; }          // float compare as in :1171-1174
    if (bestincR == -Lr || bestincR == Lr) return;
    float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
    for (int s = 1; s < 10; s++)
        if (s == bestincR + Lr) { d1 = (float)sadv[s - 1]; d2 = (float)sadv[s]; d3 = (float)sadv[s + 1]; }
    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) return;    // (the denominator is > 0: bestincR is the FIRST strict minimum, so d1 > d2 <= d3)
    float bestuR = __fmul_rn(P.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestincR), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0 && disparity < A.maxD) {
        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
        A.depth[iL] = __fdiv_rn(A.mbf, disparity);
        A.uRight[iL] = bestuR;
        A.sad[iL] = bestS;
    }
}

// single block: median of the accepted matches' SADs in (sad, iL) order, removal of everything >= 1.5 * 1.4 * median (:1203-1214)
__global__ void __launch_bounds__(1024) k_stereo_filter(StereoArgs A, int *__restrict__ n_kept) {
    __shared__ int s_cnt, s_median, s_kept;
    const int N = *A.nL;
    if (threadIdx.x == 0) { s_cnt = 0; s_median = -1; s_kept = 0; }
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) local += A.sad[i] >= 0;
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) { if (threadIdx.x == 0) *n_kept = 0; return; }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int si = A.sad[i];
        if (si < 0) continue;
        int rank = 0;                                           // position of (sad, iL) in the sorted vDistIdx
        for (int j = 0; j < N; j++) {
            const int sj = A.sad[j];
            rank += (sj >= 0) && (sj < si || (sj == si && j < i));
        }
        if (rank == cnt / 2) s_median = si;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_median);
    int kept = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int si = A.sad[i];
        if (si < 0) continue;
        if ((float)si < thDist) kept++;
        else { A.uRight[i] = -1.0f; A.depth[i] = -1.0f; }
    }
    if (kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (threadIdx.x == 0) *n_kept = s_kept;
}

}  // namespace orbx
