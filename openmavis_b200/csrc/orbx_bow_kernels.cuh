// orbx_bow_kernels.cuh — ORBmatcher::SearchByBoW on the device:
//   variant 0  SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches)          src/ORBmatcher.cc:349-666
//   variant 1  SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12)          src/ORBmatcher.cc:1006-1129
//
// Both walk two DBoW2 FeatureVectors and match, inside every vocabulary node the two share, each feature of the first list
// against the still unmatched features of the second list (best and second-best Hamming distance).  A feature belongs to exactly
// one node (FeatureVector::addFeature is called once per feature, src/Frame.cc:980-985), so the "already matched" state
// (vpMapPointMatches[realIdxF] :392 / vbMatched2[idx2] :1054) never crosses nodes: one WARP per shared node runs the node's
// sequential walk over the first list, its lanes share the scan over the second list; the nodes run in parallel.
// The rotation-histogram filter (:637-663 / :1102-1126, ComputeThreeMaxima :2537-2573) is a second, single-block kernel.
#pragma once

#include "orbx_match_kernels.cuh"

namespace orbx {

struct BowArgs {
    int variant;                       // 0 = (KeyFrame, Frame), 1 = (KeyFrame, KeyFrame)
    int npairs;                        // shared nodes
    const int *pair_a, *pair_b;        // [npairs] group index in list 1 / list 2
    const int *ptr1, *idx1, *ptr2, *idx2;
    const uint4 *desc1, *desc2;
    const uint8_t *skip1, *skip2;      // [n1] / [n2]: no usable MapPoint (or idx >= N), never null
    const float *angle1, *angle2;      // keypoint angles by feature index
    int cam_off2[5], ncam2;            // variant 0: camera blocks of F (ncam2 == 1: F.Nleft == -1)
    float nnratio;
    int check_ori;
    int *match;                        // variant 0: [n2] KF feature index per F feature; variant 1: [n1] idx2 per idx1; -1 = none
    int *claimed2;                     // [n2] second-list feature already matched (variant 1: vbMatched2; variant 0 = match != -1)
    int *events;                       // [n1 * 4] match-array index | rotation bin << 16 of every accepted match, -1 = none
    int *nmatches;
    int n1, n2;
};

struct BowBest { uint32_t k1; int d2; };   // k1 = best distance << 16 | position in the second list (first minimum), d2 = second best

__device__ __forceinline__ void bow_best_add(BowBest &b, int dist, int pos) {
    const uint32_t k = ((uint32_t)dist << 16) | (uint32_t)pos;
    if (k < b.k1) { b.d2 = min(b.d2, (int)(b.k1 >> 16)); b.k1 = k; }     // (the displaced best is the new second best)
    else b.d2 = min(b.d2, dist);
}
__device__ __forceinline__ void bow_best_merge(BowBest &a, uint32_t k1, int d2) {
    const int hi = (int)(max(a.k1, k1) >> 16);
    a.k1 = min(a.k1, k1);
    a.d2 = min(min(a.d2, d2), hi);
}

__device__ __forceinline__ int bow_rot_bin(float a1, float a2) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30.0f));
    return bin == 30 ? 0 : bin;
}

__global__ void __launch_bounds__(128) k_bow_match(const __grid_constant__ BowArgs A) {
    const int lane = threadIdx.x & 31;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= A.npairs) return;
    const int a0 = A.ptr1[A.pair_a[p]], a1 = A.ptr1[A.pair_a[p] + 1];
    const int b0 = A.ptr2[A.pair_b[p]], nb = A.ptr2[A.pair_b[p] + 1] - b0;
    const int ncls = A.variant == 0 ? A.ncam2 : 1;
    int nm = 0;
    for (int ia = a0; ia < a1; ia++) {
        const int i1 = A.idx1[ia];
        if (A.skip1[i1]) continue;                                   // warp-uniform
        const uint4 qa = A.desc1[2 * i1], qb = A.desc1[2 * i1 + 1];
        BowBest best[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { best[c].k1 = (256u << 16) | 0xFFFFu; best[c].d2 = 256; }
        for (int k = lane; k < nb; k += 32) {
            const int i2 = A.idx2[b0 + k];
            if (A.skip2[i2] || A.claimed2[i2]) continue;
            const int d = hamming256(qa, qb, A.desc2[2 * i2], A.desc2[2 * i2 + 1]);
            int cls = 0;
            if (ncls > 1) {                                           // :406-452: the camera block of realIdxF
                cls = i2 < A.cam_off2[1] ? 0 : i2 < A.cam_off2[2] ? 1 : (ncls < 4 ? -1 : i2 < A.cam_off2[3] ? 2 : i2 < A.cam_off2[4] ? 3 : -1);
                if (cls < 0) continue;
            }
#pragma unroll
            for (int c = 0; c < 4; c++) if (c == cls) bow_best_add(best[c], d, k);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c >= ncls) break;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t k1 = __shfl_xor_sync(0xFFFFFFFFu, best[c].k1, o);
                const int d2 = __shfl_xor_sync(0xFFFFFFFFu, best[c].d2, o);
                bow_best_merge(best[c], k1, d2);
            }
        }
        if (lane == 0) {
            const int d1 = (int)(best[0].k1 >> 16);
            if (A.variant == 0) {
                if (d1 <= 50) {                                                           // TH_LOW, :456
                    for (int c = 0; c < ncls; c++) {
                        const int dc = (int)(best[c].k1 >> 16);
                        // camera 0: ratio test (:457-459); cameras 1-3: `|| true` (:497-500, 536-539, 575-578)
                        const bool ok = c == 0 ? (float)dc < __fmul_rn(A.nnratio, (float)best[c].d2) : dc <= 50;
                        if (!ok) continue;
                        const int i2 = A.idx2[b0 + (int)(best[c].k1 & 0xFFFFu)];
                        A.match[i2] = i1; A.claimed2[i2] = 1; nm++;
                        if (A.check_ori) A.events[4 * i1 + c] = i2 | (bow_rot_bin(A.angle1[i1], A.angle2[i2]) << 16);
                    }
                }
            } else if (d1 < 50 && (float)d1 < __fmul_rn(A.nnratio, (float)best[0].d2)) {  // :1074-1076
                const int i2 = A.idx2[b0 + (int)(best[0].k1 & 0xFFFFu)];
                A.match[i1] = i2; A.claimed2[i2] = 1; nm++;
                if (A.check_ori) A.events[4 * i1] = i1 | (bow_rot_bin(A.angle1[i1], A.angle2[i2]) << 16);
            }
        }
        __syncwarp();
    }
    if (lane == 0 && nm) atomicAdd(A.nmatches, nm);
}

// keeps the matches of the three fullest rotation bins (ComputeThreeMaxima, src/ORBmatcher.cc:2537-2573)
__global__ void __launch_bounds__(1024) k_rot_filter(const int *__restrict__ events, int nev, int *__restrict__ match, int *__restrict__ nmatches) {
    __shared__ int s_hist[32], s_ind[3];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < nev; e += blockDim.x)
        if (events[e] >= 0) atomicAdd(&s_hist[(events[e] >> 16) & 31], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; i++) {
            const int sz = s_hist[i];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nev; e += blockDim.x) {
        const int ev = events[e];
        if (ev < 0) continue;
        const int bin = (ev >> 16) & 31;
        if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { match[ev & 0xFFFF] = -1; atomicSub(nmatches, 1); }
    }
}

}  // namespace orbx
