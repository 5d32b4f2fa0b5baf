// orbx_match_kernels.cuh — sm_100a Hamming-matching kernels (ORBmatcher / cv::BFMatcher cores).
// Descriptors are 256-bit rows (32 B, 16-byte aligned): two 128-bit loads, 8 x (xor + popc.b32) per pair.
// No tensor cores: the path is integer popcount + argmin (BASELINE.json north_star).
#pragma once

#include "orbx_common.cuh"

namespace orbx {

__global__ void k_pair_distance(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n, int *__restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------------------
// Brute-force 2-NN (cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2), src/Frame.cc:1427,1483).
// Result order = (distance asc, trainIdx asc).  Keys are packed as dist << 23 | trainIdx, so "two smallest
// keys" is exactly that order and partial results over train slices merge with plain min/max.
// One thread owns kQPT queries (registers); train rows stream through shared memory in tiles and are read
// with warp-broadcast 128-bit loads.  grid = (query blocks, train slices).
// ------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr int kKnnQPT = 2;              // queries per thread
constexpr int kKnnTile = 256;           // train rows per shared-memory tile (8 KB)
constexpr int kKeyShift = 23;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;

__device__ __forceinline__ void top2_insert(uint32_t &k0, uint32_t &k1, uint32_t k) {
    k1 = min(k1, max(k0, k));
    k0 = min(k0, k);
}

// One launch covers a GROUP of train parts (the whole train set of the single-GPU call; the own rows, or all the pulled peer rows, of
// the sharded matcher): grid.y = slices of all parts of the group, slice y belongs to the part p with first_slice[p] <= y.
// t[p] / nt[p]: rows of part p, row_off[p] = global index of its first row, slice_rows[p] = rows per slice (a multiple of kKnnTile).
struct KnnGroup { const uint4 *t[64]; int nt[64], row_off[64], slice_rows[64], first_slice[65]; int nparts; };

__global__ void __launch_bounds__(kKnnThreads) k_knn2(const uint4 *__restrict__ q, int nq, const __grid_constant__ KnnGroup G,
                                                      uint2 *__restrict__ partial) {
    int part = 0;
    while (part + 1 < G.nparts && (int)blockIdx.y >= G.first_slice[part + 1]) part++;
    const uint4 *__restrict__ t = G.t[part];
    const int nt = G.nt[part], slice_rows = G.slice_rows[part], row_off = G.row_off[part];
    __shared__ uint4 s_t[kKnnTile * 2];
    const int qbase = (blockIdx.x * kKnnThreads + threadIdx.x) * kKnnQPT;
    uint4 qa[kKnnQPT], qb[kKnnQPT];
    uint32_t k0[kKnnQPT], k1[kKnnQPT];
#pragma unroll
    for (int u = 0; u < kKnnQPT; u++) {
        const int qi = min(qbase + u, nq - 1);
        qa[u] = q[2 * qi]; qb[u] = q[2 * qi + 1];
        k0[u] = kKeyNone; k1[u] = kKeyNone;
    }
    const int t_begin = ((int)blockIdx.y - G.first_slice[part]) * slice_rows;
    const int t_end = min(nt, t_begin + slice_rows);
    for (int base = t_begin; base < t_end; base += kKnnTile) {
        const int rows = min(kKnnTile, t_end - base);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 2; i += kKnnThreads) s_t[i] = t[2 * (long long)base + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < rows; j++) {
            const uint4 ta = s_t[2 * j], tb = s_t[2 * j + 1];
            const uint32_t jj = (uint32_t)(base + j + row_off);
#pragma unroll
            for (int u = 0; u < kKnnQPT; u++) {
                const uint32_t d = (uint32_t)hamming256_csa(qa[u], qb[u], ta, tb);
                top2_insert(k0[u], k1[u], (d << kKeyShift) | jj);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kKnnQPT; u++)
        if (qbase + u < nq) partial[(long long)blockIdx.y * nq + qbase + u] = make_uint2(k0[u], k1[u]);
}

// Pull-style all-gather: copies the parts of a train set (local or peer-mapped pointers) into one contiguous local buffer with
// 128-bit loads over NVLink.  3.2 MB for 100k rows: every peer row crosses the link once instead of once per query block.
struct KnnParts { const uint4 *ptr[64]; int first_vec[65]; int nparts; };

__global__ void __launch_bounds__(256) k_pull_parts(KnnParts P, uint4 *__restrict__ dst) {
    const int total = P.first_vec[P.nparts];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int p = 0;
        while (i >= P.first_vec[p + 1]) p++;
        dst[i] = P.ptr[p][i - P.first_vec[p]];
    }
}

// Epoch flags of the sharded matcher (one 32-bit word per rank, in peer-mapped device memory): a rank publishes epoch k after its
// rows of step k are in its shared buffer; the peers' pull kernels wait for it.  Release / acquire at system scope: the flag and
// the rows cross NVLink.
__global__ void k_publish_epoch(uint32_t *flag, uint32_t epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

struct KnnFlags { const uint32_t *flag[64]; uint32_t epoch; int self; unsigned long long timeout_ns; };

// Pull with the barrier built in: thread 0 of every CTA waits until every peer part's flag has reached `epoch` (the own part needs
// no wait and is not copied: self = its index, or -1), then the CTA copies its share of the peer rows with L1-bypassing 128-bit loads.
// dst is laid out like the concatenation of ALL parts, so indices stay global.  A peer that never publishes makes the kernel trap after
// timeout_ns instead of spinning forever.
__global__ void __launch_bounds__(256) k_pull_parts_wait(KnnParts P, KnnFlags F, uint4 *__restrict__ dst) {
    if (threadIdx.x == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int p = 0; p < P.nparts; p++) {
            if (p == F.self || P.first_vec[p + 1] == P.first_vec[p]) continue;
            for (;;) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(F.flag[p]) : "memory");
                if ((int)(v - F.epoch) >= 0) break;
                __nanosleep(200);
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > F.timeout_ns) __trap();
            }
        }
    }
    __syncthreads();
    const int total = P.first_vec[P.nparts];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int p = 0;
        while (i >= P.first_vec[p + 1]) p++;
        if (p == F.self) continue;
        dst[i] = __ldcv(P.ptr[p] + (i - P.first_vec[p]));
    }
}

__global__ void k_knn2_merge(const uint2 *__restrict__ partial, int nq, int nslices, int *__restrict__ idx, int *__restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t k0 = kKeyNone, k1 = kKeyNone;
    for (int s = 0; s < nslices; s++) {
        const uint2 p = partial[(long long)s * nq + i];
        top2_insert(k0, k1, p.x);
        top2_insert(k0, k1, p.y);
    }
    // kKeyNone can only come from "no such neighbour" (a real key has dist <= 256 < 511)
    idx[2 * i] = k0 == kKeyNone ? -1 : (int)(k0 & ((1u << kKeyShift) - 1));
    dist[2 * i] = k0 == kKeyNone ? -1 : (int)(k0 >> kKeyShift);
    idx[2 * i + 1] = k1 == kKeyNone ? -1 : (int)(k1 & ((1u << kKeyShift) - 1));
    dist[2 * i + 1] = k1 == kKeyNone ? -1 : (int)(k1 >> kKeyShift);
}

// ------------------------------------------------------------------------------------------------------
// Candidate-list (CSR) matching — the inner loop of the guided matchers (src/ORBmatcher.cc:72-114 etc.):
//   if (dist < bestDist) { second = best; best = this } else if (dist < bestDist2) { second = this }
// evaluated in candidate order.  One thread per query row (lists are short: ~5-50 candidates).
// ------------------------------------------------------------------------------------------------------
__global__ void k_match_csr(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, const int *__restrict__ row_ptr,
                            const int *__restrict__ cand, int *__restrict__ best_idx, int *__restrict__ best_dist,
                            int *__restrict__ second_dist, int *__restrict__ best_pos, int *__restrict__ second_pos,
                            int *__restrict__ all_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint4 qa = q[2 * i], qb = q[2 * i + 1];
    int bd = 256, bd2 = 256, bi = -1, bp = -1, bp2 = -1;
    const int r0 = row_ptr[i], r1 = row_ptr[i + 1];
    for (int k = r0; k < r1; k++) {
        const int c = cand[k];
        const int d = hamming256(qa, qb, t[2 * c], t[2 * c + 1]);
        if (all_dist) all_dist[k] = d;
        if (d < bd) { bd2 = bd; bp2 = bp; bd = d; bi = c; bp = k - r0; }
        else if (d < bd2) { bd2 = d; bp2 = k - r0; }
    }
    if (best_idx) {
        best_idx[i] = bi; best_dist[i] = bd; second_dist[i] = bd2; best_pos[i] = bp; second_pos[i] = bp2;
    }
}

// SearchForTriangulation core (src/ORBmatcher.cc:1238-1396, bCoarse): row r = one idx1 of a shared vocabulary
// node, candidates = the idx2 list of the same node.  Keep the LAST candidate with dist <= min(TH_LOW, best so far).
__global__ void k_match_groups(const uint4 *__restrict__ d1, const uint4 *__restrict__ d2, int nrows, const int *__restrict__ row_q,
                               const int *__restrict__ row_begin, const int *__restrict__ row_len, const int *__restrict__ idx2,
                               const uint8_t *__restrict__ skip2, int th_low, int *__restrict__ match12) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int i1 = row_q[r];
    const uint4 qa = d1[2 * i1], qb = d1[2 * i1 + 1];
    int best = th_low, bi = -1;
    const int b = row_begin[r], e = b + row_len[r];
    for (int k = b; k < e; k++) {
        const int c = idx2[k];
        if (skip2[c]) continue;
        const int d = hamming256(qa, qb, d2[2 * c], d2[2 * c + 1]);
        if (d > th_low || d > best) continue;
        bi = c; best = d;
    }
    match12[i1] = bi;
}

// ------------------------------------------------------------------------------------------------------
// Frame grid (src/Frame.cc:541-582, 969-978) and window queries (src/Frame.cc:890-967), one camera.
// ------------------------------------------------------------------------------------------------------
constexpr int kGridCols = 64, kGridRows = 48;   // include/Frame.h:24-25

__device__ __forceinline__ int grid_cell_of(const orbx_kp &kp, float min_x, float min_y, float inv_w, float inv_h) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, min_x), inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, min_y), inv_h));
    if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) return -1;
    return px * kGridRows + py;
}

__global__ void k_grid_count(const orbx_kp *__restrict__ kps, int n, float min_x, float min_y, float inv_w, float inv_h,
                             int *__restrict__ cell_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = grid_cell_of(kps[i], min_x, min_y, inv_w, inv_h);
    if (c >= 0) atomicAdd(&cell_cnt[c], 1);
}

// single block: exclusive scan of the 3072 cell counts -> cell_start[3073]; resets the fill cursors
__global__ void __launch_bounds__(1024) k_grid_scan(const int *__restrict__ cell_cnt, int *__restrict__ cell_start, int *__restrict__ cursor) {
    __shared__ int s_w[32];
    const int per = (kGridCols * kGridRows) / 1024;   // 3
    int v[per], sum = 0;
#pragma unroll
    for (int k = 0; k < per; k++) { v[k] = cell_cnt[threadIdx.x * per + k]; sum += v[k]; }
    int inc = sum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_w[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += u; }
        s_w[lane] = wi - w;
    }
    __syncthreads();
    int run = s_w[wid] + inc - sum;
#pragma unroll
    for (int k = 0; k < per; k++) { cell_start[threadIdx.x * per + k] = run; cursor[threadIdx.x * per + k] = 0; run += v[k]; }
    if (threadIdx.x == 1023) cell_start[kGridCols * kGridRows] = run;
}

__global__ void k_grid_fill(const orbx_kp *__restrict__ kps, int n, float min_x, float min_y, float inv_w, float inv_h,
                            const int *__restrict__ cell_start, int *__restrict__ cursor, int *__restrict__ cell_items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = grid_cell_of(kps[i], min_x, min_y, inv_w, inv_h);
    if (c >= 0) cell_items[cell_start[c] + atomicAdd(&cursor[c], 1)] = i;
}

// insertion order inside a cell is keypoint-index order (push_back in the loop of Frame.cc:563-581)
__global__ void k_grid_sort_cells(const int *__restrict__ cell_start, int *__restrict__ cell_items) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kGridCols * kGridRows) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    for (int i = b + 1; i < e; i++) {
        const int v = cell_items[i];
        int j = i - 1;
        while (j >= b && cell_items[j] > v) { cell_items[j + 1] = cell_items[j]; j--; }
        cell_items[j + 1] = v;
    }
}

// GetFeaturesInArea for query qi: visits candidates in reference order and calls f(idx) for each accepted one.
template <class F>
__device__ __forceinline__ void grid_query(const orbx_kp *__restrict__ kps, const int *__restrict__ cell_start,
                                           const int *__restrict__ cell_items, float min_x, float min_y, float inv_w,
                                           float inv_h, float x, float y, float r, int min_level, int max_level, F f) {
    const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (c0 >= kGridCols) return;
    const int c1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (c1 < 0) return;
    const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (r0 >= kGridRows) return;
    const int r1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (r1 < 0) return;
    const bool check = (min_level > 0) || (max_level >= 0);
    // cells are stored column-major, so the cells (ix, r0..r1) of one grid column are ONE contiguous run of cell_items: two offset
    // loads per column instead of two per cell, same visiting order (iy ascending, insertion order inside a cell)
    for (int ix = c0; ix <= c1; ix++) {
        const int kb = cell_start[ix * kGridRows + r0], ke = cell_start[ix * kGridRows + r1 + 1];
        for (int k = kb; k < ke; k++) {
            const int idx = cell_items[k];
            const orbx_kp kp = kps[idx];
            if (check) {
                if (kp.octave < min_level) continue;
                if (max_level >= 0 && kp.octave > max_level) continue;
            }
            if (fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r) f(idx);
        }
    }
}

__global__ void k_grid_query_count(const orbx_kp *__restrict__ kps, const int *__restrict__ cell_start, const int *__restrict__ cell_items,
                                   float min_x, float min_y, float inv_w, float inv_h, const float *__restrict__ qx,
                                   const float *__restrict__ qy, const float *__restrict__ qr, const int *__restrict__ min_level,
                                   const int *__restrict__ max_level, int nq, int *__restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int n = 0;
    grid_query(kps, cell_start, cell_items, min_x, min_y, inv_w, inv_h, qx[i], qy[i], qr[i], min_level[i], max_level[i],
               [&](int) { n++; });
    counts[i] = n;
}

__global__ void k_grid_query_fill(const orbx_kp *__restrict__ kps, const int *__restrict__ cell_start, const int *__restrict__ cell_items,
                                  float min_x, float min_y, float inv_w, float inv_h, const float *__restrict__ qx,
                                  const float *__restrict__ qy, const float *__restrict__ qr, const int *__restrict__ min_level,
                                  const int *__restrict__ max_level, int nq, const int *__restrict__ row_ptr, int *__restrict__ cand,
                                  int cand_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int o = row_ptr[i];
    grid_query(kps, cell_start, cell_items, min_x, min_y, inv_w, inv_h, qx[i], qy[i], qr[i], min_level[i], max_level[i],
               [&](int idx) { if (o < cand_cap) cand[o] = idx; o++; });
}

// generic single-block exclusive scan (n up to a few 100k): out[n] = total
__global__ void __launch_bounds__(1024) k_excl_scan(const int *__restrict__ in, int n, int *__restrict__ out) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += u; }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < n) out[i] = carry + s_w[wid] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

// ------------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483), batched: one warp per map point.  Row i of the
// distance matrix is never materialised: the warp histograms d(i, j) over the 257 possible distances in shared
// memory and walks the cumulative counts to the median rank floor(0.5 * (N - 1)).  SURVEY.md §8(f) N4.
// ------------------------------------------------------------------------------------------------------
constexpr int kDistWarps = 4;
constexpr int kDistBins = 288;                  // >= 257, multiple of 32

__global__ void __launch_bounds__(kDistWarps * 32) k_distinctive(const uint4 *__restrict__ desc, const int *__restrict__ ptr, int ngroups,
                                                                 int *__restrict__ best) {
    __shared__ int s_hist[kDistWarps][kDistBins];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * kDistWarps + wid;
    if (g >= ngroups) return;
    const int b0 = ptr[g], N = ptr[g + 1] - b0;
    if (N <= 0) { if (lane == 0) best[g] = -1; return; }
    int *hist = s_hist[wid];
    const int k = (int)(0.5 * (double)(N - 1));                 // vDists[0.5 * (N - 1)]
    int best_median = 0x7FFFFFFF, best_idx = 0;
    for (int i = 0; i < N; i++) {
        for (int t = lane; t < kDistBins; t += 32) hist[t] = 0;
        __syncwarp();
        const uint4 qa = desc[2 * (long long)(b0 + i)], qb = desc[2 * (long long)(b0 + i) + 1];
        for (int j = lane; j < N; j += 32) atomicAdd(&hist[hamming256(qa, qb, desc[2 * (long long)(b0 + j)], desc[2 * (long long)(b0 + j) + 1])], 1);
        __syncwarp();
        // lane owns bins [9 * lane, 9 * lane + 9); the median is the first bin whose inclusive cumulative count exceeds k
        int c[kDistBins / 32], sum = 0;
#pragma unroll
        for (int t = 0; t < kDistBins / 32; t++) { c[t] = hist[(kDistBins / 32) * lane + t]; sum += c[t]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += u; }
        int run = inc - sum, med = 0x7FFFFFFF;
#pragma unroll
        for (int t = 0; t < kDistBins / 32; t++) { run += c[t]; if (run > k && med == 0x7FFFFFFF) med = (kDistBins / 32) * lane + t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) med = min(med, __shfl_xor_sync(0xFFFFFFFFu, med, o));
        if (med < best_median) { best_median = med; best_idx = i; }     // first minimum (:471-474)
        __syncwarp();
    }
    if (lane == 0) best[g] = best_idx;
}

// ------------------------------------------------------------------------------------------------------
// DBoW2 TemplatedVocabulary<FORB>::transform, per-feature tree descent
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1213-1259, FORB::distance FORB.cpp:81-101).  One warp per feature: the lanes
// take the children of the current node (k = 10 for ORBvoc), one Hamming distance each, warp arg-min with the FIRST
// minimum winning (key = distance << 8 | position), until a leaf.  SURVEY.md §8(f) N3.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vocab_transform(const int *__restrict__ child_ptr, const int *__restrict__ child,
                                                         const uint4 *__restrict__ node_desc, const int *__restrict__ word_id, int nid_level,
                                                         const uint4 *__restrict__ feat, int n, int *__restrict__ out_word,
                                                         int *__restrict__ out_node) {
    const int lane = threadIdx.x & 31;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= n) return;
    const uint4 qa = feat[2 * f], qb = feat[2 * f + 1];
    int nid = 0, cur = 0, level = 0;
    int b = child_ptr[0], e = child_ptr[1];
    while (e > b) {
        ++level;
        uint32_t best = 0xFFFFFFFFu;
        for (int k0 = b; k0 < e; k0 += 32) {          // 32 children per sweep; positions beyond 255 of a sweep cannot occur
            const int k = k0 + lane;
            uint32_t key = 0xFFFFFFFFu;
            if (k < e) {
                const int id = child[k];
                key = ((uint32_t)hamming256(qa, qb, node_desc[2 * (long long)id], node_desc[2 * (long long)id + 1]) << 20) | (uint32_t)(k - b);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, o));
            best = min(best, key);                     // earlier sweeps hold smaller positions: strict first minimum overall
        }
        cur = child[b + (int)(best & 0xFFFFFu)];
        if (level == nid_level) nid = cur;
        b = child_ptr[cur]; e = child_ptr[cur + 1];
    }
    if (lane == 0) { out_word[f] = word_id[cur]; out_node[f] = nid; }
}

// popc.b32 issue-rate microbenchmark (roofline denominator of the matching kernels)
__global__ void __launch_bounds__(256) k_popc_peak(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    uint32_t s = 0;
    for (int i = 0; i < iters; i++) {
        s += __popc(a0) + __popc(a1) + __popc(a2) + __popc(a3) + __popc(a4) + __popc(a5) + __popc(a6) + __popc(a7);
        a0 ^= s; a1 += s; a2 ^= s << 1; a3 += s >> 1; a4 ^= s << 2; a5 += s << 3; a6 ^= s >> 2; a7 += i;
    }
    if (s == 0xDEADBEEFu) out[0] = s;
}

}  // namespace orbx
