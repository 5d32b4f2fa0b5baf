// orbx_sbp_kernels.cuh — ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, ...) on the device
// (src/ORBmatcher.cc:23-340).
//
// The reference walks the map points sequentially; a keypoint claimed by an earlier point (F.mvpMapPoints[idx]
// with Observations()>0, :77-79) is invisible to later points, and a failed ratio test on camera c skips the
// remaining cameras of that point (the `continue` at :119-120 continues the MapPoint loop).  Here:
//   phase A (parallel over (point, camera) queries): window query on the 64x48 grid (src/Frame.cc:890-967) ->
//            CSR candidate lists in reference order, Hamming distance + octave per candidate;
//   phase B (one block, ordered resolve): a point is FINAL in a round iff it is the smallest unresolved point
//            on every keypoint it may read or write (its candidates and their stereo mirrors, :125-131 /
//            :195-201).  Final points of a round touch disjoint keypoints, so they are applied in parallel and
//            the result equals the sequential walk.  The smallest unresolved point is always final, so the
//            loop terminates.
//
// mode 1 = the frame-to-frame tracking matcher ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th,
// bMono) (src/ORBmatcher.cc:1985-2413) on the same machinery: radius th * sf[octave of the last keypoint] for every camera,
// level gate by bForward / bBackward, best candidate only (<= TH_HIGH), no stereo mirrors, an empty camera-0 window skips the
// point's other cameras (the `continue` at :2062), then the rotation-histogram filter (ComputeThreeMaxima :2537-2573).
//
// mode 2 = ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:895-1004): level-0 keypoints of F1 search a fixed window of F2
// around vbPrevMatched; a keypoint of F2 goes to the F1 keypoint with the smallest distance so far (vMatchedDistance, :927) and the
// previous owner is evicted (:941-944).  In the resolve phase occ[] holds vMatchedDistance and assigned[] holds vnMatches21.
#pragma once

#include "orbx_match_kernels.cuh"

namespace orbx {

struct SbpFrame {
    int ncam, n_total;
    int cam_off[5];
    const orbx_kp *kps;          // all cameras, camera-local order
    const uint4 *desc;
    float min_x[4], min_y[4], inv_w[4], inv_h[4];
    const int *cell_start[4];
    const int *cell_items[4];    // camera-local indices
    const float *scale_factors;
    const uint8_t *occupied0;    // initial occupancy (global index)
    const int *l2r, *r2l;
};

struct SbpPoints {
    int nmp;
    const float *proj_x, *proj_y, *view_cos;   // [nmp][4]
    const int *level;                          // [nmp][4]
    const uint8_t *in_view;                    // [nmp][4]
    const uint8_t *active, *has_obs;           // [nmp]
    const uint4 *desc;                         // [nmp][2]
    float th, nnratio;
    // mode 1 (SearchByProjection(CurrentFrame, LastFrame)): level[q] = octave of the last keypoint, in_view[q] = projection usable
    int mode, level_gate, check_ori, max_dist;   // level_gate: ORBX_GATE_*
    const float *angle;                        // [nmp] kpLF.angle
    const float *u_right, *proj_ur;            // mvuRight [N] / uv.x - mbf * invzc [nmp] (rectified stereo, ncam == 1), or NULL
};

struct SbpWork {
    int *counts;      // [nq]
    int *q_ptr;       // [nq+1]
    int *tmin;        // [2 N] 64-bit (round, point) stamps
    int *occ;         // [N]
    int *assigned;    // [N]
    int *resolved;    // [nmp] (unused)
    int *nmatches;    // [8]: nmatches, scratch
    int *events;      // [nmp * 4] mode 1: keypoint | rotation bin << 16 of every accepted assignment, -1 = none
    int *match12;     // [nmp] mode 2: vnMatches12
};
struct SbpResult { const int *assigned; const int *nmatches; const int *match12; };

inline size_t sbp_work_bytes(int nmp, int N) {
    return ((size_t)nmp * 4 * 3 + 2 + (size_t)N * 4 + 2 * (size_t)nmp + 8) * sizeof(int) + 256;
}
inline void sbp_carve(uint8_t *base, int nmp, int N, SbpWork *w) {
    int *p = (int *)base;
    const size_t nq = (size_t)nmp * 4;
    w->tmin = p; p += 2 * (size_t)N;       // 64-bit stamps (base is 256-byte aligned)
    w->counts = p; p += nq;
    w->q_ptr = p; p += nq + 1;
    w->occ = p; p += N;
    w->assigned = p; p += N;
    w->resolved = p; p += nmp;
    w->nmatches = p; p += 8;
    w->events = p; p += nq;
    w->match12 = p;
}
inline void sbp_result_ptrs(uint8_t *base, int nmp, int N, SbpResult *r) {
    SbpWork w; sbp_carve(base, nmp, N, &w);
    r->assigned = w.assigned; r->nmatches = w.nmatches; r->match12 = w.match12;
}

__device__ __forceinline__ bool sbp_query(const SbpFrame &F, const SbpPoints &P, int q, int &c, float &x, float &y, float &rr,
                                          int &min_lvl, int &max_lvl) {
    const int mp = q >> 2;
    c = q & 3;
    if (!P.active[mp] || !P.in_view[q]) return false;
    if (c >= 1 && F.ncam < 2) return false;     // F.Nleft == -1
    if (c >= 2 && F.ncam < 4) return false;     // F.Nsideleft == -1
    const int lvl = P.level[q];
    if (lvl < 0) return false;                  // :142 nPredictedLevel != -1 (camera 0 assumes a valid level)
    x = P.proj_x[q]; y = P.proj_y[q];
    if (P.mode == 2) { rr = P.th; min_lvl = lvl; max_lvl = lvl; return true; }    // GetFeaturesInArea(x, y, windowSize, level1, level1) :915-920
    if (P.mode == 1) {
        rr = __fmul_rn(P.th, F.scale_factors[lvl]);                            // :2040
        if (P.level_gate == ORBX_GATE_FORWARD) { min_lvl = lvl; max_lvl = -1; }            // :2044-2056
        else if (P.level_gate == ORBX_GATE_BACKWARD) { min_lvl = 0; max_lvl = lvl; }
        else if (P.level_gate == ORBX_GATE_BELOW) { min_lvl = lvl - 1; max_lvl = lvl; }      // the KeyFrame / Sim3 overloads (:759)
        else { min_lvl = lvl - 1; max_lvl = lvl + 1; }
        return true;
    }
    float r = ((double)P.view_cos[q] > 0.998) ? 2.5f : 4.0f;    // RadiusByViewingCos :342-347
    if (c == 0 && P.th != 1.0f) r = __fmul_rn(r, P.th);         // only camera 0 applies th (:51-52)
    rr = __fmul_rn(r, F.scale_factors[lvl]);
    min_lvl = lvl - 1; max_lvl = lvl;
    return true;
}

__device__ __forceinline__ int sbp_window_count(const SbpFrame &F, const SbpPoints &P, int q) {
    int c, mn, mx, n = 0; float x, y, rr;
    if (sbp_query(F, P, q, c, x, y, rr, mn, mx))
        grid_query(F.kps + F.cam_off[c], F.cell_start[c], F.cell_items[c], F.min_x[c], F.min_y[c], F.inv_w[c], F.inv_h[c], x, y, rr, mn, mx,
                   [&](int) { n++; });
    return n;
}

__global__ void k_sbp_count(SbpFrame F, SbpPoints P, int *__restrict__ counts) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nmp * 4) return;
    int n = sbp_window_count(F, P, q);
    // mode 1: `if (vIndices2.empty()) continue;` on camera 0 (:2062) leaves the point's other cameras unsearched
    if (P.mode == 1 && (q & 3) != 0 && n > 0 && sbp_window_count(F, P, q & ~3) == 0) n = 0;
    counts[q] = n;
}

// candidate record: local index | dist << 16 | octave << 25
__global__ void k_sbp_fill(SbpFrame F, SbpPoints P, const int *__restrict__ q_ptr, uint32_t *__restrict__ cand) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nmp * 4) return;
    int c, mn, mx; float x, y, rr;
    if (q_ptr[q + 1] == q_ptr[q] || !sbp_query(F, P, q, c, x, y, rr, mn, mx)) return;
    const int mp = q >> 2;
    const bool ur_check = P.mode <= 1 && c == 0 && F.ncam == 1 && P.u_right != nullptr;   // mode 0: :81-85, mode 1: :2079-2084
    const uint4 qa = P.desc[2 * mp], qb = P.desc[2 * mp + 1];
    int o = q_ptr[q];
    const orbx_kp *ck = F.kps + F.cam_off[c];
    const uint4 *cd = F.desc + 2 * (long long)F.cam_off[c];
    grid_query(ck, F.cell_start[c], F.cell_items[c], F.min_x[c], F.min_y[c], F.inv_w[c], F.inv_h[c], x, y, rr, mn, mx,
               [&](int idx) {
                   int d = hamming256(qa, qb, cd[2 * idx], cd[2 * idx + 1]);
                   if (ur_check && P.u_right[idx] > 0.0f && fabsf(__fsub_rn(P.proj_ur[mp], P.u_right[idx])) > rr) d = 511;
                   cand[o++] = (uint32_t)idx | ((uint32_t)d << 16) | ((uint32_t)(ck[idx].octave & 0x7F) << 25);
               });
}

template <class Fn>
__device__ __forceinline__ void sbp_for_touch(const SbpFrame &F, bool mirrors, const int *q_ptr, const uint32_t *cand, int mp, Fn fn) {
    for (int c = 0; c < 4; c++) {
        const int q = mp * 4 + c;
        for (int k = q_ptr[q]; k < q_ptr[q + 1]; k++) {
            const int idx = (int)(cand[k] & 0xFFFF);
            fn(idx + F.cam_off[c]);
            if (mirrors && F.ncam >= 2) {
                if (c == 0 && F.l2r[idx] != -1) fn(F.l2r[idx] + F.cam_off[1]);
                if (c == 1 && F.r2l[idx] != -1) fn(F.r2l[idx]);
            }
        }
    }
}

constexpr int kSbpTouchCap = 12;

// Resolve rounds.  Thread t owns the points t, t + 1024, ...; `mine` is its smallest unresolved one.  A round looks only at the WINDOW
// [lo, lo + 1024) of point indices, lo = the smallest unresolved point of the block: every point outside it is larger than every
// point inside, so "smallest unresolved point on each keypoint it touches" can be decided among the window's points alone — at
// most one point per thread and round.  The per-keypoint minimum is an atomicMax of (round << 32 | ~point): a new round's
// stamp outranks every older value, so the array is never cleared.
__global__ void __launch_bounds__(1024) k_sbp_resolve(SbpFrame F, SbpPoints P, SbpWork W, const uint32_t *__restrict__ cand) {
    __shared__ int s_lo;
    const int N = F.n_total, nmp = P.nmp;
    unsigned long long *tmin = reinterpret_cast<unsigned long long *>(W.tmin);
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        W.occ[g] = P.mode == 2 ? 0x7FFFFFFF : (int)F.occupied0[g]; W.assigned[g] = -1; tmin[g] = 0ull;
    }
    if (P.mode == 2)
        for (int u = threadIdx.x; u < nmp; u += blockDim.x) W.match12[u] = -1;
    if (threadIdx.x == 0) { W.nmatches[0] = 0; W.nmatches[1] = 0; }
    if (P.mode >= 1)
        for (int e = threadIdx.x; e < 4 * nmp; e += blockDim.x) W.events[e] = -1;
    // a thread's current point stays the same for many rounds: its touch set (keypoints it may read or write) is cached in shared memory
    __shared__ unsigned short s_touch[kSbpTouchCap][1024];
    int mine = threadIdx.x, nm_total = 0, loaded = -1, nt = 0, done = 0;
    for (unsigned round = 1;; round++) {
        if (threadIdx.x == 0) s_lo = 0x7FFFFFFF;
        __syncthreads();                                              // (also orders the previous round's claims before this round's reads)
        while (mine < nmp && W.q_ptr[4 * mine + 4] == W.q_ptr[4 * mine]) mine += blockDim.x;   // points without candidates
        if (mine < nmp) atomicMin(&s_lo, mine);
        if (mine < nmp && loaded != mine) {
            nt = 0;
            sbp_for_touch(F, P.mode == 0, W.q_ptr, cand, mine, [&](int g) { if (nt < kSbpTouchCap) s_touch[nt][threadIdx.x] = (unsigned short)g; nt++; });
            loaded = mine;
        }
        __syncthreads();
        const int lo = s_lo;
        if (lo == 0x7FFFFFFF) break;
        if (round > 4u * (unsigned)nmp + 16u) {   // every round resolves at least one (point, camera): more rounds = an internal error
            if (threadIdx.x == 0) W.nmatches[1] = -1;
            break;
        }
        const bool active = mine < nmp && mine < lo + (int)blockDim.x;
        const int u = mine;
        const unsigned long long key = ((unsigned long long)round << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)u);
        if (active) {
            // (mode 1: the cameras of a point are independent claim processes; a camera that is done no longer holds its keypoints back)
            auto stamp = [&](int g) {
                if (P.mode == 1) {
                    int c = 0;
                    for (int k = 1; k < F.ncam; k++) c += g >= F.cam_off[k];
                    if ((done >> c) & 1) return;
                }
                atomicMax(&tmin[g], key);
            };
            if (nt <= kSbpTouchCap) for (int j = 0; j < nt; j++) stamp(s_touch[j][threadIdx.x]);
            else sbp_for_touch(F, P.mode == 0, W.q_ptr, cand, u, stamp);
        }
        __syncthreads();
        if (active && P.mode == 1) {
            // Frame-to-frame matcher: a camera's decision depends only on its best unoccupied candidate — better ones are occupied for
            // good (occupancy only grows in this mode), worse ones do not matter — so the point may claim it as soon as it is the
            // smallest unresolved point that has this ONE keypoint in a list, instead of waiting for all its candidates.
            bool all_done = true;
            int nm = 0;
            for (int c = 0; c < 4; c++) {
                if ((done >> c) & 1) continue;
                const int q = u * 4 + c;
                const int b = W.q_ptr[q], e = W.q_ptr[q + 1];
                int bestDist = 256, bestIdx = -1;
                for (int k = b; k < e; k++) {
                    const uint32_t r = cand[k];
                    const int idx = (int)(r & 0xFFFF), dist = (int)((r >> 16) & 0x1FF);
                    if (__ldcg(&W.occ[idx + F.cam_off[c]])) continue;
                    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
                }
                if (bestDist > P.max_dist) { done |= 1 << c; continue; }          // nothing (left) to claim: final, occupancy only grows
                const int g = bestIdx + F.cam_off[c];
                if (__ldcg(&tmin[g]) != key) { all_done = false; continue; }     // an earlier point may still take it
                W.assigned[g] = u; W.occ[g] = P.has_obs[u]; nm++;                 // TH_HIGH (:2103), ORBdist (:2487), TH_LOW * ratioHamming (:767)
                if (P.check_ori) {   // :2121-2128
                    float rot = __fsub_rn(P.angle[u], F.kps[g].angle);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30.0f));
                    if (bin == 30) bin = 0;
                    W.events[q] = g | (bin << 16);
                }
                done |= 1 << c;
            }
            nm_total += nm;
            if (all_done) { mine += blockDim.x; done = 0; }
        } else if (active) {
            bool fin = true;
            if (nt <= kSbpTouchCap) { for (int j = 0; j < nt; j++) if (__ldcg(&tmin[s_touch[j][threadIdx.x]]) != key) fin = false; }   // (L2: the atomics do not update this SM's L1)
            else sbp_for_touch(F, P.mode == 0, W.q_ptr, cand, u, [&](int g) { if (__ldcg(&tmin[g]) != key) fin = false; });
            if (fin) {
            // ---- sequential body of the reference for map point u
            int nm = 0;
            for (int c = 0; c < 4; c++) {
                const int q = u * 4 + c;
                const int b = W.q_ptr[q], e = W.q_ptr[q + 1];
                if (b == e) continue;
                if (P.mode == 2) {   // :923-961 (camera 0 only)
                    int bestDist = 0x7FFFFFFF, bestDist2 = 0x7FFFFFFF, bestIdx = -1;
                    for (int k = b; k < e; k++) {
                        const uint32_t r = cand[k];
                        const int idx = (int)(r & 0xFFFF), dist = (int)((r >> 16) & 0x1FF);
                        if (W.occ[idx] <= dist) continue;                                  // vMatchedDistance[i2] <= dist
                        if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = idx; }
                        else if (dist < bestDist2) bestDist2 = dist;
                    }
                    if (bestDist <= 50 && (float)bestDist < __fmul_rn((float)bestDist2, P.nnratio)) {
                        const int old = W.assigned[bestIdx];                              // vnMatches21
                        if (old >= 0) { W.match12[old] = -1; nm--; }
                        W.match12[u] = bestIdx; W.assigned[bestIdx] = u; W.occ[bestIdx] = bestDist; nm++;
                        if (P.check_ori) {
                            float rot = __fsub_rn(P.angle[u], F.kps[bestIdx].angle);
                            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                            int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30.0f));
                            if (bin == 30) bin = 0;
                            W.events[q] = u | (bin << 16);
                        }
                    }
                    break;
                }
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
                for (int k = b; k < e; k++) {
                    const uint32_t r = cand[k];
                    const int idx = (int)(r & 0xFFFF), dist = (int)((r >> 16) & 0x1FF), oct = (int)(r >> 25);
                    if (W.occ[idx + F.cam_off[c]]) continue;
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
                    else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
                }
                if (bestDist <= 100) {   // TH_HIGH
                    if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(P.nnratio, (float)bestDist2)) break;
                    const int g = bestIdx + F.cam_off[c];
                    const int ho = P.has_obs[u];
                    if (c == 0) {
                        W.assigned[g] = u; W.occ[g] = ho;
                        if (F.ncam >= 2 && F.l2r[bestIdx] != -1) {
                            const int gr = F.l2r[bestIdx] + F.cam_off[1];
                            W.assigned[gr] = u; W.occ[gr] = ho; nm++;
                        }
                        nm++;
                    } else if (c == 1) {
                        if (F.r2l[bestIdx] != -1) { const int gl = F.r2l[bestIdx]; W.assigned[gl] = u; W.occ[gl] = ho; nm++; }
                        W.assigned[g] = u; W.occ[g] = ho; nm++;
                    } else {
                        W.assigned[g] = u; W.occ[g] = ho; nm++;
                    }
                }
            }
            nm_total += nm;
            mine += blockDim.x;
            }
        }
    }
    if (nm_total) atomicAdd(&W.nmatches[0], nm_total);
    __syncthreads();
    if (P.mode == 0 || !P.check_ori) return;
    // ---- rotation consistency (:2392-2410): keep the assignments of the three fullest histogram bins
    __shared__ int s_hist[32], s_ind[3];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * nmp; e += blockDim.x)
        if (W.events[e] >= 0) atomicAdd(&s_hist[(W.events[e] >> 16) & 31], 1);
    __syncthreads();
    if (threadIdx.x == 0) {   // ComputeThreeMaxima (:2537-2573)
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
    for (int e = threadIdx.x; e < 4 * nmp; e += blockDim.x) {
        const int ev = W.events[e];
        if (ev < 0) continue;
        const int bin = (ev >> 16) & 31;
        if (bin == s_ind[0] || bin == s_ind[1] || bin == s_ind[2]) continue;
        if (P.mode == 2) {   // :982-985: only matches that were not evicted meanwhile
            if (W.match12[ev & 0xFFFF] >= 0) { W.match12[ev & 0xFFFF] = -1; atomicSub(&W.nmatches[0], 1); }
        } else { W.assigned[ev & 0xFFFF] = -2; atomicSub(&W.nmatches[0], 1); }   // -2 = claimed, then cleared (the reference writes NULL, :2404); the host reports -1
    }
}

// ------------------------------------------------------------------------------------------------------
// Window search without claims: the search loop shared by Fuse (src/ORBmatcher.cc:1549-1601, 1719-1738) and SearchBySim3
// (:1837-1861, 1911-1935): KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:771-834) window, level gate, optional reprojection gate
// (:1571-1592), first minimum of the Hamming distance.  One thread per query.
// ------------------------------------------------------------------------------------------------------
struct WindowArgs {
    const orbx_kp *kps; const uint4 *desc;
    const int *cell_start, *cell_items;
    float min_x, min_y, inv_w, inv_h;
    const uint8_t *skip;                 // [n] or null
    const float *u_right;                // [n] or null (mvuRight, stereo gate)
    const float *inv_sigma2;             // [nlevels] or null (no reprojection gate)
    int nq;
    const uint8_t *valid;
    const float *qx, *qy, *q_ur, *radius;
    const int *min_level, *max_level;
    const uint4 *q_desc;
    int *best_idx, *best_dist;
};

__global__ void __launch_bounds__(128) k_window_best(const __grid_constant__ WindowArgs A) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.nq) return;
    int best = 256, bi = -1;
    if (A.valid[q]) {
        const float x = A.qx[q], y = A.qy[q];
        const int mn = A.min_level[q], mx = A.max_level[q];
        const uint4 qa = A.q_desc[2 * q], qb = A.q_desc[2 * q + 1];
        grid_query(A.kps, A.cell_start, A.cell_items, A.min_x, A.min_y, A.inv_w, A.inv_h, x, y, A.radius[q], -1, -1, [&](int idx) {
            if (A.skip && A.skip[idx]) return;
            const orbx_kp kp = A.kps[idx];
            if (kp.octave < mn || kp.octave > mx) return;
            if (A.inv_sigma2) {
                const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                if (A.u_right && A.u_right[idx] >= 0.0f) {
                    const float er = __fsub_rn(A.q_ur[q], A.u_right[idx]);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, A.inv_sigma2[kp.octave]) > 7.8) return;
                } else if ((double)__fmul_rn(e2, A.inv_sigma2[kp.octave]) > 5.99) return;
            }
            const int d = hamming256(qa, qb, A.desc[2 * idx], A.desc[2 * idx + 1]);
            if (d < best) { best = d; bi = idx; }
        });
    }
    A.best_idx[q] = bi; A.best_dist[q] = best;
}

// host driver (orbx_match.cu)
template <class DBufT>
inline int sbp_run(SbpFrame F, SbpPoints P, uint8_t *work, DBufT &candbuf, cudaStream_t s, int *launches) {
    SbpWork W;
    sbp_carve(work, P.nmp, F.n_total, &W);
    const int nq = P.nmp * 4;
    k_sbp_count<<<(nq + 127) / 128, 128, 0, s>>>(F, P, W.counts);
    k_excl_scan<<<1, 1024, 0, s>>>(W.counts, nq, W.q_ptr);
    int total = 0;
    ORBX_CUDA(cudaMemcpyAsync(&total, W.q_ptr + nq, sizeof(int), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    ORBX_CUDA(candbuf.ensure((size_t)(total + 1) * sizeof(uint32_t)));
    uint32_t *cand = (uint32_t *)candbuf.p;
    k_sbp_fill<<<(nq + 127) / 128, 128, 0, s>>>(F, P, W.q_ptr, cand);
    k_sbp_resolve<<<1, 1024, 0, s>>>(F, P, W, cand);
    *launches += 4;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

}  // namespace orbx
