// orbx_extract_kernels.cuh — sm_100a kernels of the ORB extraction path.
//
// Data layout in HBM: every pyramid level of every image of a batch is a "plane" of
// rows = h+38 lines of `pitch` bytes (pitch % 128 == 0); the level interior starts at row 19, column 32, so
// the 19-px BORDER_REFLECT_101 frame of mvImagePyramid (src/ORBextractor.cc:1077-1101) is physically
// present and interior rows are 16-byte aligned.  Planes of one level are contiguous over the batch
// (level-major), so one launch per stage covers all cameras x frames.
#pragma once

#include "math_core.cuh"
#include "octree_core.cuh"
#include "orbx_common.cuh"

namespace orbx {

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};   // src/ORBextractor.cc:399-413

__device__ __forceinline__ const uint8_t *plane_ptr(const uint8_t *arena, const LevelGeom &L, int img) {
    return arena + L.plane_off + (long long)img * L.plane_bytes;
}
__device__ __forceinline__ uint8_t *plane_ptr(uint8_t *arena, const LevelGeom &L, int img) {
    return arena + L.plane_off + (long long)img * L.plane_bytes;
}

// ------------------------------------------------------------------------------------------------------
// K1a  level 0: copyMakeBorder(image, temp, 19.., BORDER_REFLECT_101)  (src/ORBextractor.cc:1094-1101)
// one thread = 16 output bytes of one plane row (one 128-bit store)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_level0(const uint8_t *__restrict__ img, long long img_bytes, int img_stride,
                                                uint8_t *__restrict__ pyr, LevelGeom L) {
    const int chunks = (kXO + L.w + kEdge + 15) >> 4;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * L.rows) return;
    const int row = idx / chunks, c0 = (idx - row * chunks) << 4;
    const int b = blockIdx.y;
    const uint8_t *src = img + (long long)b * img_bytes;
    const int iy = orbx_reflect101(row - kYO, L.h);
    const uint8_t *srow = src + (long long)iy * img_stride;
    uint4 v;
    const int x0 = c0 - kXO;
    if (x0 >= 0 && x0 + 16 <= L.w && ((reinterpret_cast<uintptr_t>(srow + x0) & 15) == 0)) {
        v = __ldg(reinterpret_cast<const uint4 *>(srow + x0));
    } else {
        uint32_t wds[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t wv = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int x = x0 + 4 * k + j;
                x = max(-kEdge, min(L.w - 1 + kEdge, x));
                wv |= (uint32_t)__ldg(srow + orbx_reflect101(x, L.w)) << (8 * j);
            }
            wds[k] = wv;
        }
        v = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    *reinterpret_cast<uint4 *>(plane_ptr(pyr, L, b) + (long long)row * L.pitch + c0) = v;
}

// ------------------------------------------------------------------------------------------------------
// K1b  level l>0: resize(level l-1 -> level l, INTER_LINEAR) + copyMakeBorder(REFLECT_101|ISOLATED)
// (src/ORBextractor.cc:1083-1092).  The border is produced by evaluating the resize at the reflected
// interior coordinate, so a level is finished by a single pass and only the previous interior is read.
// Fixed-point arithmetic of OpenCV's 8-bit linear resize, exact.
// one thread = 4 output pixels (one 32-bit store)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(uint8_t *__restrict__ pyr, LevelGeom S, LevelGeom D,
                                                const ResizeTap *__restrict__ taps) {
    const int chunks = (kXO + D.w + kEdge + 3) >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * D.rows) return;
    const int row = idx / chunks, c0 = (idx - row * chunks) << 2;
    const int b = blockIdx.y;
    const uint8_t *src = plane_ptr((const uint8_t *)pyr, S, b) + (long long)kYO * S.pitch + kXO;
    const int iy = orbx_reflect101(row - kYO, D.h);
    const ResizeTap ty = taps[D.tap_y + iy];
    const uint8_t *s0 = src + (long long)ty.s0 * S.pitch;
    const uint8_t *s1 = src + (long long)ty.s1 * S.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int x = c0 + j - kXO;
        x = max(-kEdge, min(D.w - 1 + kEdge, x));
        const ResizeTap tx = taps[D.tap_x + orbx_reflect101(x, D.w)];
        const int r0 = (int)s0[tx.s0] * tx.a0 + (int)s0[tx.s1] * tx.a1;
        const int r1 = (int)s1[tx.s0] * tx.a0 + (int)s1[tx.s1] * tx.a1;
        const int v = (((ty.a0 * (r0 >> 4)) >> 16) + ((ty.a1 * (r1 >> 4)) >> 16) + 2) >> 2;
        out |= (uint32_t)(v & 0xFF) << (8 * j);
    }
    *reinterpret_cast<uint32_t *>(plane_ptr(pyr, D, b) + (long long)row * D.pitch + c0) = out;
}

// ------------------------------------------------------------------------------------------------------
// K5  GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every level (src/ORBextractor.cc:1035-1036).
// Q8.8 kernel {18,34,48,56,48,34,18}: horizontal u8->u16, vertical u16->u32, (v + 2^15) >> 16 — exact.
// The clone's REFLECT_101 edge is already materialised by the plane's 19-px frame, so the tile load needs
// no index arithmetic.  Tile = 128 x 32 outputs; source tile staged in shared memory with 32-bit loads.
// ------------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 32;
constexpr int kBlurSW = kBlurTW + 8;          // source columns x0-4 .. x0+131
constexpr int kBlurSH = kBlurTH + 6;

struct BlurTiles { int base[kMaxLevels + 1]; };

__global__ void __launch_bounds__(256) k_blur(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, Geom G,
                                              BlurTiles T) {
    __shared__ uint32_t s_src[kBlurSH][kBlurSW / 4];
    __shared__ uint16_t s_h[kBlurSH][kBlurTW + 8];
    int lvl = 0;
    while (lvl + 1 < G.nlevels && (int)blockIdx.x >= T.base[lvl + 1]) lvl++;
    const LevelGeom &L = G.lv[lvl];
    const int tiles_x = (L.w + kBlurTW - 1) / kBlurTW;
    const int t = blockIdx.x - T.base[lvl];
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
    const int b = blockIdx.y;
    const uint8_t *src = plane_ptr(pyr, L, b);
    // source tile: plane rows (kYO + y0 - 3 ...), plane columns (kXO + x0 - 4 ...), 4-byte aligned
    for (int i = threadIdx.x; i < kBlurSH * (kBlurSW / 4); i += blockDim.x) {
        const int r = i / (kBlurSW / 4), c = i - r * (kBlurSW / 4);
        const int pr = min(kYO + y0 - 3 + r, L.rows - 1);
        const int pc = kXO + x0 - 4 + 4 * c;
        uint32_t v = 0;
        if (pc + 4 <= L.pitch) v = __ldg(reinterpret_cast<const uint32_t *>(src + (long long)pr * L.pitch + pc));
        s_src[r][c] = v;
    }
    __syncthreads();
    // horizontal: 4 adjacent outputs per item from 3 words (bytes x-4..x+7, taps x-3..x+6)
    for (int i = threadIdx.x; i < kBlurSH * (kBlurTW / 4); i += blockDim.x) {
        const int r = i / (kBlurTW / 4), c = i - r * (kBlurTW / 4);
        const uint32_t w0 = s_src[r][c], w1 = s_src[r][c + 1], w2 = s_src[r][c + 2];
        int p[12];
#pragma unroll
        for (int k = 0; k < 4; k++) { p[k] = (w0 >> (8 * k)) & 0xFF; p[4 + k] = (w1 >> (8 * k)) & 0xFF; p[8 + k] = (w2 >> (8 * k)) & 0xFF; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int acc = 18 * (p[k + 1] + p[k + 7]) + 34 * (p[k + 2] + p[k + 6]) + 48 * (p[k + 3] + p[k + 5]) + 56 * p[k + 4];
            s_h[r][4 * c + k] = (uint16_t)acc;
        }
    }
    __syncthreads();
    uint8_t *dst = plane_ptr(blur, L, b);
    for (int i = threadIdx.x; i < kBlurTH * (kBlurTW / 4); i += blockDim.x) {
        const int r = i / (kBlurTW / 4), c = i - r * (kBlurTW / 4);
        const int y = y0 + r, x = x0 + 4 * c;
        if (y >= L.h || x >= L.w) continue;
        uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int kw = (k == 0 || k == 6) ? 18 : (k == 1 || k == 5) ? 34 : (k == 2 || k == 4) ? 48 : 56;
            const uint2 hv = *reinterpret_cast<const uint2 *>(&s_h[r + k][4 * c]);
            acc[0] += kw * (hv.x & 0xFFFF); acc[1] += kw * (hv.x >> 16);
            acc[2] += kw * (hv.y & 0xFFFF); acc[3] += kw * (hv.y >> 16);
        }
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= ((acc[k] + 32768u) >> 16) << (8 * k);
        uint8_t *d = dst + (long long)(kYO + y) * L.pitch + kXO + x;
        if (x + 4 <= L.w) *reinterpret_cast<uint32_t *>(d) = out;
        else for (int k = 0; k < L.w - x; k++) d[k] = (uint8_t)(out >> (8 * k));
    }
}

// ------------------------------------------------------------------------------------------------------
// K2  per-cell FAST-9-16 with non-max suppression and the iniThFAST -> minThFAST retry
// (src/ORBextractor.cc:726-794; cv::FAST semantics in SURVEY.md §8a-a3).  One block per (cell, image).
// The cell ROI and its score map live in shared memory; candidates are emitted in row-major order with a
// ballot/popc prefix so the output order equals OpenCV's.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_rect(const LevelGeom &L, int i, int j, int &x0, int &y0, int &cw, int &ch) {
    const int ini_y = kFastBorder + i * L.h_cell;
    if (ini_y >= L.max_by - 3) return false;
    const int max_y = min(ini_y + L.h_cell + 6, L.max_by);
    const int ini_x = kFastBorder + j * L.w_cell;
    if (ini_x >= L.max_bx - 6) return false;
    const int max_x = min(ini_x + L.w_cell + 6, L.max_bx);
    x0 = ini_x; y0 = ini_y; cw = max_x - ini_x; ch = max_y - ini_y;
    return true;
}

// score of a pixel that is known to be a corner: max over the 16 arcs of 9 of min(d) (dark) or min(-d)
// (bright), minus 1 == cornerScore<16>() for every threshold the pixel passes.
__device__ __forceinline__ int arc9_max_min(const int *d) {
    int m3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    int best = -1000;
#pragma unroll
    for (int k = 0; k < 16; k++) best = max(best, min(min(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15]));
    return best;
}

__global__ void __launch_bounds__(256) k_fast(const uint8_t *__restrict__ pyr, Geom G,
                                              const orbx_image_params *__restrict__ params, uint32_t *__restrict__ cells,
                                              int *__restrict__ cell_cnt, int tile_pitch) {
    extern __shared__ uint8_t smem[];
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.y;
    int lvl = 0;
    while (lvl + 1 < G.nlevels && (int)blockIdx.x >= G.lv[lvl + 1].cell_base) lvl++;
    const LevelGeom &L = G.lv[lvl];
    const int cell = blockIdx.x - L.cell_base;
    const int ci = cell / L.n_cols, cj = cell - ci * L.n_cols;
    int *cnt_out = cell_cnt + (long long)b * G.total_cells + blockIdx.x;
    int x0, y0, cw, ch;
    if (!cell_rect(L, ci, cj, x0, y0, cw, ch) || cw < 7 || ch < 7) {
        if (threadIdx.x == 0) *cnt_out = 0;
        return;
    }
    uint8_t *tile = smem;                                  // [ch][tile_pitch]
    uint8_t *score = smem + (size_t)tile_pitch * (L.h_cell + 6);   // same shape
    const uint8_t *src = plane_ptr(pyr, L, b) + (long long)(kYO + y0) * L.pitch + kXO + x0;
    for (int i = threadIdx.x; i < cw * ch; i += blockDim.x) {
        const int r = i / cw, c = i - r * cw;
        tile[r * tile_pitch + c] = __ldg(src + (long long)r * L.pitch + c);
    }
    const int dw = cw - 6, dh = ch - 6;
    uint32_t *out = cells + (long long)b * G.cand_per_image + L.cand_off + (long long)cell * L.cell_cap;
    const int off_x = cj * L.w_cell, off_y = ci * L.h_cell;
    const orbx_image_params prm = params[b];
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int t = min(max(pass == 0 ? prm.ini_th_fast : prm.min_th_fast, 0), 255);
        for (int i = threadIdx.x; i < cw * ch; i += blockDim.x) score[(i / cw) * tile_pitch + (i % cw)] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
            const int r = i / dw + 3, c = i - (i / dw) * dw + 3;
            const uint8_t *p = tile + r * tile_pitch + c;
            const int v = p[0];
            // high-speed reject (necessary condition for an arc of 9): each diametral pair needs a pass
            const int lo = v - t, hi = v + t;
            const int a0 = p[3 * tile_pitch], a8 = p[-3 * tile_pitch];
            if (!((a0 < lo) || (a8 < lo) || (a0 > hi) || (a8 > hi))) continue;
            const int a4 = p[3], a12 = p[-3];
            if (!((a4 < lo) || (a12 < lo) || (a4 > hi) || (a12 > hi))) continue;
            int d[16];
            d[0] = v - a0; d[4] = v - a4; d[8] = v - a8; d[12] = v - a12;
            d[1] = v - p[3 * tile_pitch + 1]; d[2] = v - p[2 * tile_pitch + 2]; d[3] = v - p[tile_pitch + 3];
            d[5] = v - p[-tile_pitch + 3]; d[6] = v - p[-2 * tile_pitch + 2]; d[7] = v - p[-3 * tile_pitch + 1];
            d[9] = v - p[-3 * tile_pitch - 1]; d[10] = v - p[-2 * tile_pitch - 2]; d[11] = v - p[-tile_pitch - 3];
            d[13] = v - p[tile_pitch - 3]; d[14] = v - p[2 * tile_pitch - 2]; d[15] = v - p[3 * tile_pitch - 1];
            uint32_t md = 0, mb = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) { md |= (d[k] > t ? 1u : 0u) << k; mb |= (d[k] < -t ? 1u : 0u) << k; }
            md |= md << 16; mb |= mb << 16;
            uint32_t xd = md & (md >> 1); xd &= xd >> 2; xd &= xd >> 4; xd &= md >> 8;
            uint32_t xb = mb & (mb >> 1); xb &= xb >> 2; xb &= xb >> 4; xb &= mb >> 8;
            int sc = 0;
            if (xd & 0xFFFF) {
                sc = arc9_max_min(d) - 1;
            } else if (xb & 0xFFFF) {
#pragma unroll
                for (int k = 0; k < 16; k++) d[k] = -d[k];
                sc = arc9_max_min(d) - 1;
            }
            if (sc > 0) score[r * tile_pitch + c] = (uint8_t)sc;
        }
        __syncthreads();
        // 3x3 strict non-max suppression + ordered compaction
        total = 0;
        for (int s = 0; s < dw * dh; s += blockDim.x) {
            const int i = s + threadIdx.x;
            bool keep = false;
            int r = 0, c = 0, sc = 0;
            if (i < dw * dh) {
                r = i / dw + 3; c = i - (i / dw) * dw + 3;
                const uint8_t *q = score + r * tile_pitch + c;
                sc = q[0];
                keep = sc > 0 && sc > q[-1] && sc > q[1] && sc > q[-tile_pitch - 1] && sc > q[-tile_pitch] &&
                       sc > q[-tile_pitch + 1] && sc > q[tile_pitch - 1] && sc > q[tile_pitch] && sc > q[tile_pitch + 1];
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            if (lane == 0) s_warp[wid] = __popc(bal);
            __syncthreads();
            int before = 0, all = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { const int n = s_warp[k]; all += n; if (k < wid) before += n; }
            if (keep) {
                const int pos = total + before + __popc(bal & ((1u << lane) - 1));
                if (pos < L.cell_cap) out[pos] = cand_pack(c + off_x, r + off_y, sc);
            }
            total += all;
            __syncthreads();
        }
        if (total > 0) break;
    }
    if (threadIdx.x == 0) *cnt_out = min(total, L.cell_cap);
    (void)s_base;
}

// ------------------------------------------------------------------------------------------------------
// K3  DistributeOctTree per (level, image): gathers the cell lists in vToDistributeKeys order, runs the
// array quadtree (octree_core.cuh) and writes the level's keypoints (level coordinates, +16 border).
// ------------------------------------------------------------------------------------------------------
struct DevCtx {
    int tid, nt;
    int *total;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ int atomic_add(int *p, int v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_max(unsigned *p, unsigned v) { atomicMax(p, v); }
    __device__ int excl_scan(int *a, int n) {
        if (tid < 32) {
            int carry = 0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + tid;
                const int v = i < n ? a[i] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (tid >= o) inc += u; }
                if (i < n) a[i] = carry + inc - v;
                carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            if (tid == 0) *total = carry;
        }
        __syncthreads();
        return *total;
    }
};

__global__ void __launch_bounds__(512) k_octree(Geom G, const int *__restrict__ nfeat_lvl, const uint32_t *__restrict__ cells,
                                                const int *__restrict__ cell_cnt, int *__restrict__ cell_off,
                                                uint32_t *__restrict__ lin, int *__restrict__ lin_n,
                                                unsigned short *__restrict__ node, int *__restrict__ sel_tmp,
                                                uint32_t *__restrict__ lvl_kp, int *__restrict__ lvl_n, int list_cap) {
    extern __shared__ __align__(16) unsigned char oct_smem[];
    __shared__ int s_total;
    const int lvl = blockIdx.x, b = blockIdx.y;
    const LevelGeom &L = G.lv[lvl];
    DevCtx cx; cx.tid = threadIdx.x; cx.nt = blockDim.x; cx.total = &s_total;
    const int ncell = L.n_cols * L.n_rows;
    const int *cc = cell_cnt + (long long)b * G.total_cells + L.cell_base;
    int *co = cell_off + (long long)b * G.total_cells + L.cell_base;
    for (int i = cx.tid; i < ncell; i += cx.nt) co[i] = cc[i];
    __syncthreads();
    const int n = cx.excl_scan(co, ncell);
    __syncthreads();
    const uint32_t *cs = cells + (long long)b * G.cand_per_image + L.cand_off;
    uint32_t *ln = lin + (long long)b * G.cand_per_image + L.cand_off;
    unsigned short *nd = node + (long long)b * G.cand_per_image + L.cand_off;
    const int lane = cx.tid & 31, wid = cx.tid >> 5, nw = cx.nt >> 5;
    for (int c = wid; c < ncell; c += nw) {
        const int k = cc[c], o = co[c];
        for (int j = lane; j < k; j += 32) ln[o + j] = cs[(long long)c * L.cell_cap + j];
    }
    __syncthreads();
    const int N = nfeat_lvl[b * G.nlevels + lvl];
    OctWork w;
    oct_carve(w, oct_smem, list_cap);
    int *sel = sel_tmp + (long long)b * G.kp_per_image + L.kp_off;
    int k = 0;
    if (n > 0 && oct_list_cap(N, L.n_ini) <= list_cap && oct_list_cap(N, L.n_ini) <= L.kp_cap + 1)
        k = oct_distribute(cx, w, ln, nd, n, L.max_bx - kFastBorder, L.max_by - kFastBorder, L.n_ini, L.h_x, N, sel);
    __syncthreads();
    k = min(k, L.kp_cap);
    uint32_t *kp = lvl_kp + (long long)b * G.kp_per_image + L.kp_off;
    for (int p = cx.tid; p < k; p += cx.nt) {
        const uint32_t c = ln[sel[p]];
        kp[p] = cand_pack(cand_x(c) + kFastBorder, cand_y(c) + kFastBorder, cand_score(c));
    }
    if (cx.tid == 0) { lvl_n[b * G.nlevels + lvl] = k; lin_n[b * G.nlevels + lvl] = n; }
}

// ------------------------------------------------------------------------------------------------------
// K7a  output slots (src/ORBextractor.cc:1024-1068): levels ascending, keypoints in list order; a keypoint
// whose scaled x lies in [lap0, lap1] takes slot stereoIndex-- (from n-1 down), the others monoIndex++.
// One block per image.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slots(Geom G, const orbx_image_params *__restrict__ params,
                                               const uint32_t *__restrict__ lvl_kp, const int *__restrict__ lvl_n,
                                               int *__restrict__ slot, int *__restrict__ n_kp, int *__restrict__ n_mono) {
    __shared__ int s_warp[8];
    const int b = blockIdx.x;
    const orbx_image_params prm = params[b];
    int n = 0;
    for (int l = 0; l < G.nlevels; l++) n += lvl_n[b * G.nlevels + l];
    int mono = 0, stereo = 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int l = 0; l < G.nlevels; l++) {
        const LevelGeom &L = G.lv[l];
        const int k = lvl_n[b * G.nlevels + l];
        const uint32_t *kp = lvl_kp + (long long)b * G.kp_per_image + L.kp_off;
        int *sl = slot + (long long)b * G.kp_per_image + L.kp_off;
        for (int s = 0; s < k; s += blockDim.x) {
            const int i = s + threadIdx.x;
            bool valid = i < k, st = false;
            if (valid) {
                float x = (float)cand_x(kp[i]);
                if (l != 0) x = __fmul_rn(x, L.scale);
                st = x >= (float)prm.lap0 && x <= (float)prm.lap1;
            }
            const unsigned bs = __ballot_sync(0xFFFFFFFFu, valid && st), bm = __ballot_sync(0xFFFFFFFFu, valid && !st);
            if (lane == 0) s_warp[wid] = __popc(bs) | (__popc(bm) << 16);
            __syncthreads();
            int bs_before = 0, bm_before = 0, bs_all = 0, bm_all = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int v = s_warp[q];
                bs_all += v & 0xFFFF; bm_all += v >> 16;
                if (q < wid) { bs_before += v & 0xFFFF; bm_before += v >> 16; }
            }
            const unsigned lm = (1u << lane) - 1;
            if (valid) sl[i] = st ? (n - 1 - (stereo + bs_before + __popc(bs & lm))) : (mono + bm_before + __popc(bm & lm));
            stereo += bs_all; mono += bm_all;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) { n_kp[b] = n; n_mono[b] = mono; }
}

// ------------------------------------------------------------------------------------------------------
// K4 + K6 + K7b  one warp per keypoint: IC_Angle on the unblurred level (src/ORBextractor.cc:19-43),
// 256-bit rBRIEF on the blurred level (src/ORBextractor.cc:46-90), final KeyPoint record + descriptor row.
// Orientation: lane = column u of the 31-wide patch, loop over the 31 rows; int32 moments reduced with
// shuffles (order-independent, exact).  Descriptor: lane = output byte, 16 rotated samples per lane.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_orient_desc(const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur, Geom G,
                                                     const signed char *__restrict__ pattern, const uint32_t *__restrict__ lvl_kp,
                                                     const int *__restrict__ lvl_n, const int *__restrict__ slot,
                                                     float *__restrict__ lvl_angle, orbx_kp *__restrict__ out_kp,
                                                     uint8_t *__restrict__ out_desc) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int b = blockIdx.y;
    if (warp >= G.kp_per_image) return;
    int lvl = 0;
    while (lvl + 1 < G.nlevels && warp >= G.lv[lvl + 1].kp_off) lvl++;
    const LevelGeom &L = G.lv[lvl];
    const int i = warp - L.kp_off;
    if (i >= lvl_n[b * G.nlevels + lvl]) return;
    const uint32_t kp = lvl_kp[(long long)b * G.kp_per_image + warp];
    const int x = cand_x(kp), y = cand_y(kp);
    // ---- orientation
    const uint8_t *c = plane_ptr(pyr, L, b) + (long long)(kYO + y) * L.pitch + kXO + x;
    const int u = lane - kHalfPatch;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int au = u < 0 ? -u : u;
#pragma unroll 1
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
            if (au <= c_umax[v < 0 ? -v : v]) {
                const int val = __ldg(c + (long long)v * L.pitch + u);
                m10 += u * val; m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
    const float angle = orbx_fast_atan2((float)m01, (float)m10);
    // ---- descriptor
    const float rad = __fmul_rn(angle, 0.0174532924f);   // (float)(CV_PI/180.f)
    float sn, cs;
    orbx_sincosf(rad, &sn, &cs);
    const uint8_t *cb = plane_ptr(blur, L, b) + (long long)(kYO + y) * L.pitch + kXO + x;
    const signed char *pp = pattern + lane * 32;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int tv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)pp[(2 * k + e) * 2], py = (float)pp[(2 * k + e) * 2 + 1];
            float fy, fx;
            if (G.fma_mode) { fy = __fmaf_rn(px, sn, __fmul_rn(py, cs)); fx = __fmaf_rn(px, cs, -__fmul_rn(py, sn)); }
            else { fy = __fadd_rn(__fmul_rn(px, sn), __fmul_rn(py, cs)); fx = __fsub_rn(__fmul_rn(px, cs), __fmul_rn(py, sn)); }
            tv[e] = __ldg(cb + (long long)__float2int_rn(fy) * L.pitch + __float2int_rn(fx));
        }
        val |= (tv[0] < tv[1] ? 1 : 0) << k;
    }
    const int s = slot[(long long)b * G.kp_per_image + warp];
    out_desc[((long long)b * G.out_cap + s) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        lvl_angle[(long long)b * G.kp_per_image + warp] = angle;
        orbx_kp o;
        o.x = (float)x; o.y = (float)y;
        if (lvl != 0) { o.x = __fmul_rn(o.x, L.scale); o.y = __fmul_rn(o.y, L.scale); }
        o.size = L.patch; o.angle = angle; o.response = (float)cand_score(kp); o.octave = lvl; o.class_id = -1;
        out_kp[(long long)b * G.out_cap + s] = o;
    }
}

}  // namespace orbx
