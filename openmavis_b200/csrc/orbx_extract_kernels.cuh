// orbx_extract_kernels.cuh — sm_100a kernels of the ORB extraction path.
//
// Data layout in HBM: every pyramid level of every image of a batch is a "plane" of
// rows = h+38 lines of `pitch` bytes (pitch % 128 == 0); the level interior starts at row 19, column 32, so
// the 19-px BORDER_REFLECT_101 frame of mvImagePyramid (src/ORBextractor.cc:1077-1101) is physically
// present and interior rows are 16-byte aligned.  Planes of one level are contiguous over the batch
// (level-major), so one launch per stage covers all cameras x frames.
#pragma once

#include <limits.h>

#include "math_core.cuh"
#include "octree_core.cuh"
#include "orbx_common.cuh"
#include "tma.cuh"

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
// one thread = 16 output bytes (one 128-bit store) of kL0Rows consecutive plane rows: the loads of all rows are
// issued before the first store.
// ------------------------------------------------------------------------------------------------------
constexpr int kL0Rows = 4;   // (8 rows per thread measured slower: pyramid 0.548 vs 0.537 ms)

__device__ __forceinline__ uint4 level0_chunk(const uint8_t *__restrict__ srow, int x0, int w) {
    if (x0 >= 0 && x0 + 16 <= w && ((reinterpret_cast<uintptr_t>(srow + x0) & 15) == 0))
        return __ldg(reinterpret_cast<const uint4 *>(srow + x0));
    uint32_t wds[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t wv = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int x = x0 + 4 * k + j;
            x = max(-kEdge, min(w - 1 + kEdge, x));           // plane padding outside the frame repeats the frame's edge
            x = x < 0 ? -x : (x >= w ? 2 * (w - 1) - x : x);   // BORDER_REFLECT_101, one reflection (w > 19)
            wv |= (uint32_t)__ldg(srow + x) << (8 * j);
        }
        wds[k] = wv;
    }
    return make_uint4(wds[0], wds[1], wds[2], wds[3]);
}

__global__ void __launch_bounds__(256) k_level0(const uint8_t *__restrict__ img, long long img_bytes, int img_stride,
                                                uint8_t *__restrict__ pyr, LevelGeom L, int img0) {
    const int chunks = (kXO + L.w + kEdge + 15) >> 4;
    const int row_groups = (L.rows + kL0Rows - 1) / kL0Rows;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * row_groups) return;
    // Work order: first every chunk that lies inside the image row (one aligned 128-bit load), then the few frame chunks
    // per row (byte gathers with reflection) — otherwise every warp would contain a few frame chunks and run both paths.
    const int c_int0 = kXO >> 4, n_int = L.w >> 4, n_frame = chunks - n_int;
    int rg, c;
    if (idx < n_int * row_groups) { rg = idx / n_int; c = c_int0 + (idx - rg * n_int); }
    else { const int j = idx - n_int * row_groups; rg = j / n_frame; c = j - rg * n_frame; if (c >= c_int0) c += n_int; }
    const int c0 = c << 4;
    const int b = blockIdx.y + img0;
    const uint8_t *src = img + (long long)blockIdx.y * img_bytes;
    uint4 v[kL0Rows];
#pragma unroll
    for (int k = 0; k < kL0Rows; k++) {
        const int row = min(rg * kL0Rows + k, L.rows - 1);
        int y = row - kYO;
        y = y < 0 ? -y : (y >= L.h ? 2 * (L.h - 1) - y : y);
        v[k] = level0_chunk(src + (long long)y * img_stride, c0 - kXO, L.w);
    }
    uint8_t *dst = plane_ptr(pyr, L, b) + c0;
#pragma unroll
    for (int k = 0; k < kL0Rows; k++) {
        const int row = rg * kL0Rows + k;
        if (row < L.rows) *reinterpret_cast<uint4 *>(dst + (long long)row * L.pitch) = v[k];
    }
}

// ------------------------------------------------------------------------------------------------------
// K1b  level l>0: resize(level l-1 -> level l, INTER_LINEAR) + copyMakeBorder(REFLECT_101|ISOLATED)
// (src/ORBextractor.cc:1083-1092).  Fixed-point arithmetic of OpenCV's 8-bit linear resize, exact
// (SURVEY.md §8a-a2):  r[sy][d] = S[sy][sx]*a0 + S[sy][sx+1]*a1 ;  dst = (((b0*(r[sy]>>4))>>16) + ((b1*(r[sy+1]>>4))>>16) + 2) >> 2.
// CTA tile = 64 x 32 interior outputs.  The source rows the tile needs (about 1.2 x 32 + 2) are staged in shared
// memory with 16-byte loads; phase H resizes every staged row horizontally ONCE into a u16 buffer (r >> 4 < 2^15);
// phase V combines two buffered rows per output, 4 pixels per thread, one 32-bit store.  The 19-px REFLECT_101 frame is
// written by the CTA that produced the mirrored interior pixel (extra row / byte stores on edge tiles only).
// ------------------------------------------------------------------------------------------------------
// (Round 2 measured a register-marching variant — 128 x 64 tiles, a lane holds 4 output columns, the horizontal pass of a source row
//  lives in two ping-pong register sets and is reused by the next output row, no shared-memory round trip, equal-sized tiles: 104 M
//  instead of 123 M warp-instructions on level 1 of the bench batch, but 67 % instead of 80 % issue utilisation — 145 vs 143 us.  Not kept.)
constexpr int kRszTW = 64, kRszTH = 64;
constexpr int kRszThreads = 256;
// dynamic shared memory: staged source [src_rows][src_pitch] u8, horizontal sums [src_rows][kRszTW] u32, y taps [kRszTH], row mirrors [kRszTH];
// src_rows / src_pitch are the maxima over all tiles of all levels, computed by the host from the tap tables
__host__ __device__ inline size_t resize_smem_bytes(int src_rows, int src_pitch) {
    return (size_t)src_rows * src_pitch + (size_t)src_rows * kRszTW * 4 + kRszTH * sizeof(int4) + kRszTH * sizeof(int);
}

// REFLECT_101 image of interior row y in the top / bottom frame, or INT_MIN (levels are at least 68 rows high: at most one)
__device__ __forceinline__ int resize_row_mirror(int y, int h) {
    if (y >= 1 && y <= kEdge) return -y;
    if (y <= h - 2 && y >= h - 1 - kEdge) return 2 * (h - 1) - y;
    return INT_MIN;
}

// ((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2 with b already shifted left by 16
__device__ __forceinline__ uint32_t resize_vtap(uint32_t b0, uint32_t h0, uint32_t b1, uint32_t h1) {
    uint32_t a, r;
    asm("mad.hi.u32 %0, %1, %2, 2;" : "=r"(a) : "r"(b0), "r"(h0));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(b1), "r"(h1), "r"(a));
    return r >> 2;
}

// source level of one resize launch as a TMA tensor: dims (pitch bytes, plane rows, images), box = (src_pitch, src_rows, 1)
struct ResizeMap { CUtensorMap m; int use_tma; };

// grid = (tiles_x, tiles_y, images)
__global__ void __launch_bounds__(kRszThreads) k_resize(uint8_t *__restrict__ pyr, LevelGeom S, LevelGeom D,
                                                        const ResizeTap *__restrict__ taps, int src_rows, int src_pitch,
                                                        const __grid_constant__ ResizeMap map, int img0) {
    extern __shared__ __align__(128) uint8_t rsz_smem[];
    uint8_t *s_src = rsz_smem;
    uint32_t *s_h = reinterpret_cast<uint32_t *>(rsz_smem + (size_t)src_rows * src_pitch);
    int4 *s_ty = reinterpret_cast<int4 *>(s_h + (size_t)src_rows * kRszTW);   // {row0 offset, row1 offset, b0 << 16, b1 << 16}
    const int ox0 = blockIdx.x * kRszTW, oy0 = blockIdx.y * kRszTH;
    const int ow = min(kRszTW, D.w - ox0), oh = min(kRszTH, D.h - oy0);
    const int b = blockIdx.z + img0;
    const int tid = threadIdx.x;
    const ResizeTap *tapx = taps + D.tap_x + ox0, *tapy = taps + D.tap_y + oy0;
    const int sy_lo = tapy[0].s0, sy_hi = tapy[oh - 1].s1;
    const int sx_lo = tapx[0].s0, sx_hi = tapx[ow - 1].s1;
    const int sx_al = sx_lo & ~15;                              // staged columns start 16-byte aligned
    const int nrow = sy_hi - sy_lo + 1, nvec = (sx_hi - sx_al + 16) >> 4;
    int *s_ym = reinterpret_cast<int *>(s_ty + kRszTH);        // row of the tile row's image in the top / bottom frame, or INT_MIN
    if (tid < oh) {
        const ResizeTap t = tapy[tid];
        s_ty[tid] = make_int4((t.s0 - sy_lo) * kRszTW, (t.s1 - sy_lo) * kRszTW, (int)t.a0 << 16, (int)t.a1 << 16);
        s_ym[tid] = resize_row_mirror(oy0 + tid, D.h);
    }
    __shared__ __align__(8) uint64_t s_bar;
    if (map.use_tma) {
        // ONE TMA tensor copy stages the source tile (box = src_pitch x src_rows from (sx_al, sy_lo); the part of the box beyond the
        // rows / columns this tile needs is harmless, beyond the plane it is zero-filled)
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)(src_rows * src_pitch));
            tma_load_3d(s_src, &map.m, kXO + sx_al, kYO + sy_lo, b, &s_bar);
        }
        if (tid < 32) mbar_wait(&s_bar, 0);
    } else {   // 16-byte loads
        const uint8_t *src = plane_ptr((const uint8_t *)pyr, S, b) + (long long)(kYO + sy_lo) * S.pitch + kXO + sx_al;
        for (int c = tid & 7; c < nvec; c += 8)                 // 8 vectors (128 B) per row pass
            for (int r = tid >> 3; r < nrow; r += kRszThreads / 8)
                *reinterpret_cast<uint4 *>(s_src + r * src_pitch + 16 * c) = __ldg(reinterpret_cast<const uint4 *>(src + (long long)r * S.pitch + 16 * c));
    }
    __syncthreads();
    {   // phase H: thread = two adjacent output columns, rows strided by 8:  r[sy][d] >> 4.  Both columns' taps lie within 8 staged
        // bytes from the word that holds the first one (scale factors up to 3): two 32-bit loads per row, the byte pairs are
        // funnel-shifted into place and each column is ONE IDP.2A with its (a0, a1) coefficient pair.
        const int c2 = 2 * (tid & 31), rg = tid >> 5;
        if (c2 < ow) {
            const ResizeTap t0 = tapx[c2], t1 = tapx[min(c2 + 1, ow - 1)];
            const int o0 = (t0.s0 - sx_al) & 3, o1 = t1.s0 - ((t0.s0 - sx_al) & ~3) - sx_al;      // byte offsets from the first word
            const uint32_t k0 = (uint32_t)(unsigned short)t0.a0 | ((uint32_t)(unsigned short)t0.a1 << 16);
            const uint32_t k1 = (uint32_t)(unsigned short)t1.a0 | ((uint32_t)(unsigned short)t1.a1 << 16);
            uint32_t *hp = s_h + rg * kRszTW + c2;
            constexpr int kRows = kRszThreads / 32;
            if (o1 <= 6) {
                const uint8_t *sr = s_src + rg * src_pitch + ((t0.s0 - sx_al) & ~3);
                const int sh0 = 8 * o0, sh1a = min(8 * o1, 32), sh1b = max(8 * o1 - 32, 0);
                const int step = kRows * src_pitch;
#pragma unroll 4
                for (int r = rg; r < nrow; r += kRows) {
                    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(sr), w1 = *reinterpret_cast<const uint32_t *>(sr + 4);
                    const uint32_t p0 = __funnelshift_r(w0, w1, sh0), p1 = __funnelshift_rc(w0, w1, sh1a) >> sh1b;
                    *reinterpret_cast<uint2 *>(hp) = make_uint2(__dp2a_lo(k0, p0, 0u) >> 4, __dp2a_lo(k1, p1, 0u) >> 4);
                    sr += step; hp += kRows * kRszTW;
                }
            } else {   // wider tap spacing: byte loads
                const uint8_t *sa = s_src + rg * src_pitch + (t0.s0 - sx_al), *sb = s_src + rg * src_pitch + (t1.s0 - sx_al);
                const int da = t0.s1 - t0.s0, db = t1.s1 - t1.s0;
                for (int r = rg; r < nrow; r += kRows) {
                    hp[0] = (uint32_t)(((int)sa[0] * t0.a0 + (int)sa[da] * t0.a1) >> 4);
                    hp[1] = (uint32_t)(((int)sb[0] * t1.a0 + (int)sb[db] * t1.a1) >> 4);
                    sa += kRows * src_pitch; sb += kRows * src_pitch; hp += kRows * kRszTW;
                }
            }
        }
    }
    __syncthreads();
    // phase V: thread = 4 adjacent output columns, rows strided by 16:
    //   dst = (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2, with (b * h) >> 16 == umulhi(b << 16, h)
    // The staged source is dead now: its space keeps the output tile for the frame pass below.
    uint32_t *s_out = reinterpret_cast<uint32_t *>(s_src);     // [kRszTH][kRszTW / 4] words
    const bool y_band = oy0 <= kEdge || oy0 + oh - 1 >= D.h - 1 - kEdge;       // tile touches rows mirrored into the frame
    const bool x_band = ox0 <= kEdge || ox0 + ow - 1 >= D.w - 1 - kEdge;
    uint8_t *dplane = plane_ptr(pyr, D, b) + kXO;
    const int c4 = 4 * (tid & 15);
    if (c4 < ow) {
        const int x = ox0 + c4;
        const int nvalid = min(4, D.w - x);
        uint8_t *drow = dplane + (long long)(kYO + oy0 + (tid >> 4)) * D.pitch + x;
        const long long rstep = (long long)(kRszThreads / 16) * D.pitch;
        for (int r = tid >> 4; r < oh; r += kRszThreads / 16, drow += rstep) {
            const int4 t = s_ty[r];
            const uint4 h0 = *reinterpret_cast<const uint4 *>(s_h + t.x + c4);
            const uint4 h1 = *reinterpret_cast<const uint4 *>(s_h + t.y + c4);
            // (the rounding constant rides on the first multiply-high's addend: two IMAD.HI + one shift per pixel)
            const uint32_t p0 = resize_vtap((uint32_t)t.z, h0.x, (uint32_t)t.w, h1.x);
            const uint32_t p1 = resize_vtap((uint32_t)t.z, h0.y, (uint32_t)t.w, h1.y);
            const uint32_t p2 = resize_vtap((uint32_t)t.z, h0.z, (uint32_t)t.w, h1.z);
            const uint32_t p3 = resize_vtap((uint32_t)t.z, h0.w, (uint32_t)t.w, h1.w);
            const uint32_t out = __byte_perm(__byte_perm(p0, p1, 0x0040), __byte_perm(p2, p3, 0x0040), 0x5410);
            if (nvalid == 4) *reinterpret_cast<uint32_t *>(drow) = out;
            else for (int k = 0; k < nvalid; k++) drow[k] = (uint8_t)(out >> (8 * k));
            if (x_band) s_out[r * (kRszTW / 4) + (c4 >> 2)] = out;
            if (y_band) {   // this row's image in the top / bottom frame
                const int ym = s_ym[r];
                if (ym != INT_MIN) {
                    uint8_t *mrow = dplane + (long long)(kYO + ym) * D.pitch + x;
                    if (nvalid == 4) *reinterpret_cast<uint32_t *>(mrow) = out;
                    else for (int k = 0; k < nvalid; k++) mrow[k] = (uint8_t)(out >> (8 * k));
                }
            }
        }
    }
    if (!x_band) return;
    // frame pass (tiles at the left / right edge only): interior columns 1..19 -> -1..-19, w-2..w-20 -> w..w+18, for every row
    // of the tile and its own image in the top / bottom frame
    __syncthreads();
    if (ox0 == 0) {
        // left frame, one aligned word per item: columns -4j-4 .. -4j-1 are the interior columns 4j+4 .. 4j+1 — the bytes of two
        // adjacent words of the output tile in reverse order (ONE PRMT).  Column -20 lies outside the frame: word 4 is written as
        // one byte + one half-word.  (Round 2: was one byte store per frame pixel, 13 % of the kernel's instructions.)
        for (int i = tid; i < oh * 5; i += kRszThreads) {
            const int r = i / 5, j = i - r * 5;
            const uint32_t wv = __byte_perm(s_out[r * (kRszTW / 4) + j], s_out[r * (kRszTW / 4) + j + 1], 0x1234);
            const int ym = s_ym[r];
            uint8_t *d0 = dplane + (long long)(kYO + oy0 + r) * D.pitch - 4 * j - 4;
            uint8_t *d1 = dplane + (long long)(kYO + (ym != INT_MIN ? ym : oy0 + r)) * D.pitch - 4 * j - 4;   // (no mirror: the same word again)
            if (j < 4) {
                *reinterpret_cast<uint32_t *>(d0) = wv;
                *reinterpret_cast<uint32_t *>(d1) = wv;
            } else {
                d0[1] = (uint8_t)(wv >> 8); *reinterpret_cast<unsigned short *>(d0 + 2) = (unsigned short)(wv >> 16);
                d1[1] = (uint8_t)(wv >> 8); *reinterpret_cast<unsigned short *>(d1 + 2) = (unsigned short)(wv >> 16);
            }
        }
    }
    if (ox0 + ow - 1 >= D.w - 1 - kEdge) {
        // right frame: byte stores (the mirrored columns may straddle the last two tiles; each tile writes the pixels it holds)
        const uint8_t *ob = reinterpret_cast<const uint8_t *>(s_out);
        for (int i = tid; i < oh * kEdge; i += kRszThreads) {
            const int r = i / kEdge, k = i - r * kEdge + 1;                 // k = 1..19
            const int xx = D.w - 1 - k - ox0;                               // mirrored interior column inside the tile
            if (xx < 0 || xx >= ow) continue;
            const uint8_t pv = ob[r * kRszTW + xx];
            const int ym = s_ym[r];
            dplane[(long long)(kYO + oy0 + r) * D.pitch + D.w - 1 + k] = pv;
            if (ym != INT_MIN) dplane[(long long)(kYO + ym) * D.pitch + D.w - 1 + k] = pv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K5  GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every level (src/ORBextractor.cc:1035-1036).
// Q8.8 kernel {18,34,48,56,48,34,18}: horizontal u8->u16, vertical u16->u32, (v + 2^15) >> 16 — exact.
// The clone's REFLECT_101 edge is already materialised by the plane's 19-px frame, so the tile load needs
// no index arithmetic.
// CTA tile = 128 x 128 outputs staged once in shared memory (16-byte loads); each warp owns a 128 x 32 strip and
// MARCHES down it: a lane holds 4 adjacent columns, computes the horizontal sums of one source row with
// 10 IDP.4A (the taps are shifted inside the weight words) and keeps the last rows of sums in registers as packed row pairs, so
// the vertical pass is 3 IDP.2A + 1 multiply-add per pixel with no intermediate buffer.  One 32-bit store per lane and row.
// ------------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 128;
constexpr int kBlurStrip = 32;                      // rows marched by one warp
constexpr int kBlurWarps = kBlurTH / kBlurStrip;
constexpr int kBlurSP = kBlurTW + 32;               // staged columns x0-16 .. x0+144
constexpr int kBlurSR = kBlurTH + 6;
constexpr size_t kBlurSmem = (size_t)kBlurSP * kBlurSR;

struct BlurTiles { int base[kMaxLevels + 1]; };
// one tensor map per level over the pyramid arena: dims (pitch bytes, plane rows, images), box = the staged blur tile
struct BlurMaps { CUtensorMap m[kMaxLevels]; int use_tma; };

__constant__ uint32_t kBlurW[13] = {0x30221200u, 0x12223038u, 0x22120000u, 0x22303830u, 0x00000012u, 0x12000000u, 0x30383022u,
                                    0x00001222u, 0x38302212u, 0x00122230u, 0x2212u, 0x3830u, 0x2230u};

__global__ void __launch_bounds__(kBlurWarps * 32) k_blur(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, Geom G,
                                                          const int *__restrict__ tile_info, const __grid_constant__ BlurMaps maps, int img0) {
    extern __shared__ __align__(128) uint8_t s_src[];   // [kBlurSR][kBlurSP]
    const int info = __ldg(tile_info + blockIdx.x);            // level | tile row << 4 | tile column << 16
    const int lvl = info & 15;
    const LevelGeom &L = G.lv[lvl];
    const int x0 = (info >> 16) * kBlurTW, y0 = ((info >> 4) & 0xFFF) * kBlurTH;
    const int b = blockIdx.y + img0;
    const uint8_t *src = plane_ptr(pyr, L, b);
    // stage plane rows kYO + y0 - 3 .., plane columns kXO + x0 - 16 .. (16-byte aligned)
    const int rows_needed = min(kBlurSR, L.h + 3 - (y0 - 3));
    __shared__ __align__(8) uint64_t s_bar;
    if (maps.use_tma) {
        // ONE TMA tensor copy stages the whole 160 x 134 tile (UTMALDG); rows / columns beyond the plane are zero-filled
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)(kBlurSP * kBlurSR));
            tma_load_3d(s_src, &maps.m[lvl], kXO + x0 - 16, kYO + y0 - 3, b, &s_bar);
        }
        if (threadIdx.x < 32) mbar_wait(&s_bar, 0);   // one warp polls the mbarrier; the others park on the hardware barrier below
    } else {   // 16-byte cp.async (LDGSTS) copies of the source rows, clipped to the plane pitch at the right edge
        const int pc = kXO + x0 - 16;
        const int nvec = min(kBlurSP, L.pitch - pc) >> 4;
        const uint8_t *g = src + (long long)(kYO + y0 - 3) * L.pitch + pc;
        const uint32_t d = smem_u32(s_src);
        const int c = threadIdx.x & 15, r0 = threadIdx.x >> 4;           // 16 vector columns (10 used) x 8 rows per pass
        if (c < nvec)
            for (int r = r0; r < rows_needed; r += (kBlurWarps * 32) / 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + r * kBlurSP + 16 * c), "l"(g + (long long)r * L.pitch + 16 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ys = y0 + wid * kBlurStrip;              // first output row of this warp's strip
    const int xb = x0 + 4 * lane;
    if (ys >= L.h || xb >= L.w) return;
    const uint8_t *sp = s_src + (wid * kBlurStrip) * kBlurSP + 12 + 4 * lane;
    uint8_t *dst = plane_ptr(blur, L, b) + (long long)(kYO + ys) * L.pitch + kXO + xb;
    const int nrows = min(kBlurStrip, L.h - ys);       // output rows of the strip
    unsigned long long drow = (unsigned long long)__cvta_generic_to_global(dst);
    int nrows_k = nrows;
    asm("" : "+r"(nrows_k));
    const long long pitch = L.pitch;
    // register window: the horizontal sums (<= 65280, 16 bits) of the last rows, packed in pairs  pr[j] = h[row - 1] | h[row] << 16,
    // so that the vertical pass is 3 IDP.2A (two taps each) + 1 multiply-add per pixel
    uint32_t pr[6][4], hprev[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        hprev[k] = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) pr[j][k] = 0;
    }
    // Fully unrolled with no control flow around the register window (rows below the strip compute on stale shared memory
    // and are simply not stored), so the window is pure register renaming.
#pragma unroll
    for (int r = 0; r < kBlurStrip + 6; r++) {
        const uint32_t wl = *reinterpret_cast<const uint32_t *>(sp + r * kBlurSP);
        const uint32_t wc = *reinterpret_cast<const uint32_t *>(sp + r * kBlurSP + 4);
        const uint32_t wr = *reinterpret_cast<const uint32_t *>(sp + r * kBlurSP + 8);
        uint32_t h[4];
        // pixel k needs bytes k+1 .. k+7 of the 12-byte window (wl, wc, wr): instead of shifting the pixels into place, the
        // kernel {18,34,48,56,48,34,18} is shifted inside the weight words (10 IDP.4A, no funnel shifts)
        // (the weight words are constant-bank operands: as immediates ptxas re-materialised them into a handful of uniform registers
        //  nine times per row)
        h[0] = __dp4a(wl, kBlurW[0], __dp4a(wc, kBlurW[1], 0u));
        h[1] = __dp4a(wl, kBlurW[2], __dp4a(wc, kBlurW[3], __dp4a(wr, kBlurW[4], 0u)));
        h[2] = __dp4a(wl, kBlurW[5], __dp4a(wc, kBlurW[6], __dp4a(wr, kBlurW[7], 0u)));
        h[3] = __dp4a(wc, kBlurW[8], __dp4a(wr, kBlurW[9], 0u));
        if (r >= 6) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++)   // rows r-6 .. r with weights 18 34 48 56 48 34 18: pairs (r-6, r-5), (r-4, r-3), (r-2, r-1), and row r
                v[k] = __dp2a_lo(pr[1][k], kBlurW[10], __dp2a_lo(pr[3][k], kBlurW[11], __dp2a_lo(pr[5][k], kBlurW[12], 18u * h[k] + 32768u)));
            const uint32_t lo = __byte_perm(v[0], v[1], 0x0062), hi = __byte_perm(v[2], v[3], 0x0062);
            // (a running row pointer the compiler cannot re-derive: it rebuilt dst + (r - 6) * pitch from the kernel parameters
            //  for every row otherwise, ten instructions per store)
            if (r - 6 < nrows_k) asm volatile("st.global.u32 [%0], %1;" ::"l"(drow), "r"(__byte_perm(lo, hi, 0x5410)) : "memory");
            drow += pitch;
            asm("" : "+l"(drow));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {   // pr[j] holds the pair that ends at row r - 5 + j
#pragma unroll
            for (int j = 0; j < 5; j++) pr[j][k] = pr[j + 1][k];
            pr[5][k] = __byte_perm(hprev[k], h[k], 0x5410);   // hprev | h << 16 (both < 2^16) as ONE PRMT
            hprev[k] = h[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K2  per-cell FAST-9-16 with non-max suppression and the iniThFAST -> minThFAST retry
// (src/ORBextractor.cc:726-794; cv::FAST semantics in SURVEY.md §8a-a3).
//
// Formulation.  With ring values r[0..15] around a centre v,
//     A = min over the 16 arcs of 9 of (max of the arc),   B = max over the arcs of (min of the arc),
//     S = max(v - A, B - v) - 1
// is exactly cornerScore<16>() of every pixel that is a corner, and a pixel is a corner at threshold t iff
// S >= t (an arc of 9 with v - r > t exists iff v - A > t; the bright case is symmetric; the two cases
// exclude each other because two arcs of 9 overlap on a 16-ring).  S does not depend on t, and cv::FAST's
// 3x3 non-max suppression keeps p iff S(p) is strictly greater than the scores of its 8 neighbours inside
// the tested area of the SAME cell ROI (outside it the score buffer is 0), so
//     keypoints of a cell at threshold t  =  { p : S(p) >= t, S(p) > S(q) for the neighbours q of the cell }.
// The tested areas of the cells tile the level without gaps or overlap, which makes the work regular:
// one CTA = one tile of up to G adjacent cells of a cell row, staged once in shared memory, one warp per cell.
//   0. the tile (4 cells) is staged once (one TMA tensor copy); from then on every WARP works alone on its own cell — no CTA barriers, no
//      shared counters, no atomics except the keep bitmap words two neighbouring cells share;
//   1. quick reject with byte-SIMD (VABSDIFF4): a corner needs |v - r| > t on one pixel of each of the four opposite pairs of the
//      even ring positions (fast_quick8).  One lane marches down a strip of 13 rows of one 4-pixel word column with the column in
//      registers (no index arithmetic in the loop); the survivor flags of a lane end up in two registers (7 + 6 rows);
//   2. survivors are compacted into the warp's queue (one warp scan per work item, then a branch-free walk over the flag positions
//      with predicated stores) and scored two at a time with packed u16x2 3-input min/max (VIMNMX3.U16x2): no divergence, no
//      per-threshold corner test;
//   3. strict 3x3 maximum inside the cell over the warp's corner list -> keep bitmap;
//   4. a cell without a keypoint repeats 1-3 with minThFAST (only its own warp, rare);
//   5. emission in row-major order (the order of cv::FAST): a lane takes 64 bits of the keep bitmap, one warp scan.
// ------------------------------------------------------------------------------------------------------
constexpr int kFastWarps = 4;                 // warps per CTA = cells per tile (one warp per cell)
constexpr int kFastThreads = kFastWarps * 32;
constexpr int kFastStrip = 13;                // rows a lane marches through per work item (one 4-pixel word wide): flag registers of 7 + 6 rows.
                                              // (Cells are 35-42 rows high: three 13-row strips march 39 rows where three 14-row strips marched 42 —
                                              //  0.660 -> 0.649 ms per 320 images; 12-row strips need a fourth strip: 0.778 ms)
constexpr int kFastHalf = 7;                  // rows per flag register (byte = pixel, bit = row)
constexpr int kFastMaxStrips = 13;            // strips with host-chosen start rows (FastTile::sr); taller cells use equal strips
constexpr int kFastQueue = 512;               // per-warp survivor queue (u16 entries): < 64 left + up to 448 new (a flag register with more
                                              // survivors is appended in four parts)
constexpr int kFastMaxG = kFastWarps;
constexpr int kFastCorners = 256;             // corner list of one cell (per warp); a cell with more corners falls back to a sweep

// One FAST tile (CTA) of an image: up to kFastMaxG adjacent cells of one cell row; built on the host once per handle.
struct __align__(16) FastTile {
    int lvl, ncell, wc, cell_cap;      // level, cells in this tile, nominal cell width, candidate slots per cell
    int tx0, ty0, X, Y;                // tested area in level coordinates: x in [tx0, tx0 + X), y in [ty0, ty0 + Y)
    int pc0, prow0, c0, cell0;         // staged tile: plane column (16-B aligned) and plane row of its corner; smem column of tx0; first cell index
    int nstr, rs, cand0_lo, cand0_hi;  // strips of rs <= kFastStrip rows; slot offset of the first cell's candidate slots in an image's arena
    long long plane_off, plane_bytes;  // (cp.async fallback) byte offset of image 0's plane, bytes per plane
    int pitch;
    unsigned short sr[kFastMaxStrips + 1];   // nstr <= kFastMaxStrips: first row of strip s (sr[nstr] = Y); a strip owns the rows up to the next
                                             // strip's first row.  The host picks the start rows so that the strips of one warp iteration
                                             // fall into different shared-memory banks (fast_plan_strips)
};
static_assert(sizeof(FastTile) == 112, "FastTile layout");

// shared memory of one CTA (ip = row pitch in bytes, a template constant so that every row offset is an immediate):
//   img [(ymax + 6)][ip]               level pixels, row 0 = tile row -3, column 0 = plane column pc0 (16-byte aligned)
//   sc  [(ymax + 2)][ip]               scores, same column indexing, rows 0 and Y+1 stay zero
//   per warp: q [kFastQueue] u16 survivor queue (entries = byte offsets into img), cl [kFastCorners] u16 corner list,
//   kb [kb_words] u32 keep bitmap of the warp's cell (bit = row * cell width + column)
__host__ __device__ inline int fast_zero_bytes(int ip, int ymax) { return ((ymax + 2) * ip + 15) & ~15; }
__host__ __device__ inline int fast_warp_bytes(int kb_words) { return kb_words * 4 + (kFastQueue + kFastCorners) * 2; }
__host__ __device__ inline size_t fast_smem_bytes(int ip, int ymax, int kb_words) {
    return (size_t)(ymax + 6) * ip + fast_zero_bytes(ip, ymax) + (size_t)kFastWarps * fast_warp_bytes(kb_words);
}

// exact x / d for 0 <= x < 65536, 8 <= d: (x + 0.5) / d is at least 0.5 / d away from an integer, far above fp32 error
__device__ __forceinline__ int small_div(int x, float inv_d) { return (int)(((float)x + 0.5f) * inv_d); }

// Host side: strips of a tile.  A lane marches down rs rows of one word column; item f of a cell = (strip f / nw, word f % nw).  The
// count / height of the strips is chosen for lane utilisation, their first rows for bank spread: the lanes of one warp iteration read
// word (row * P + word) of the tile, so two strips whose first rows differ by d rows are d * P banks apart.
inline void fast_plan_strips(int Y, int wc, int ip, int *nstr_out, int *rs_out, unsigned short *sr) {
    const int P = ip / 4;
    int nws[4];
    for (int off = 0; off < 4; off++) nws[off] = (off + wc + 3) / 4;   // words a cell row spans at each alignment of its first pixel
    const int nw = nws[1];                                             // bank spread is evaluated for a common alignment
    // greedy start rows for n strips of rs rows: strip s starts at the row in [prev + 1, prev + rs] (coverage of the rest permitting)
    // that collides least with the lanes already placed in the same 32-lane group; returns the number of colliding lanes
    auto place = [&](int n, int rs, unsigned short *st) {
        int total = 0;
        st[0] = 0;
        for (int s = 1; s < n; s++) {
            // (every strip lies inside the Y rows, so the march never reads past the staged tile; a strip may own no row at all)
            const int lo = std::max((int)st[s - 1], Y - (n - s) * rs), hi = std::min(st[s - 1] + rs, Y - rs);
            int best_r = hi, best_c = 1 << 30;
            for (int r = hi; r >= lo; r--) {                           // (late starts first: fewer rows tested twice)
                int c = 0;
                for (int w = 0; w < nw; w++)
                    for (int s2 = 0; s2 < s; s2++)
                        for (int w2 = 0; w2 < nw; w2++)
                            if (((s2 * nw + w2) >> 5) == ((s * nw + w) >> 5) && ((st[s2] * P + w2 - r * P - w) & 31) == 0) c++;
                if (c < best_c) { best_c = c; best_r = r; }
            }
            st[s] = (unsigned short)best_r;
            total += best_c;
        }
        st[n] = (unsigned short)Y;
        return total;
    };
    const int nmin = (Y + kFastStrip - 1) / kFastStrip;
    int best_n = nmin, best_rs = (Y + nmin - 1) / nmin;
    double best_cost = 1e30;
    unsigned short st[kFastMaxStrips + 1];
    for (int s = 0; s <= kFastMaxStrips; s++) sr[s] = 0;
    for (int n = nmin; n <= std::min(std::max(nmin, Y), kFastMaxStrips); n++)
        for (int rs = Y >= kFastStrip ? kFastStrip : (Y + n - 1) / n; rs <= std::min(Y, kFastStrip); rs++) {   // (full-height strips run the
                                                                                                               //  branch-free variant of the quick test)
            double cost = 0;
            for (int off = 0; off < 4; off++) cost += 0.25 * ((n * nws[off] + 31) / 32) * (rs * 36.0 + 130.0);   // instructions per cell pass
            cost += 0.5 * place(n, rs, st) * (3 * rs + 14);                                                        // extra LDS wavefronts
            if (cost < best_cost) { best_cost = cost; best_n = n; best_rs = rs; memcpy(sr, st, sizeof(st)); }
        }
    *nstr_out = best_n; *rs_out = best_rs;   // (more than kFastMaxStrips strips: equal strips s * rs, no table)
}

// Shared memory is addressed through 32-bit shared-window addresses held in registers the compiler cannot re-derive (sa_keep): with
// generic pointers it rebuilt the per-warp base addresses (S2R / cluster-window LEA chains) inside every inner loop.  *_ro loads are for
// the read-only image tile (free to schedule); the others are ordered (volatile, memory clobber): scores, queue, lists, keep bits.
__device__ __forceinline__ uint32_t sa_keep(uint32_t a) { asm volatile("" : "+r"(a)::"memory"); return a; }
__device__ __forceinline__ uint32_t lds32_ro(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8_ro(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128_zero(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }
__device__ __forceinline__ void reds_or32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_add64(uint32_t a, unsigned long long v) { asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// Quick reject of one work item: rows [0, rs) of one word column (4 pixels), col = shared address of the word in the strip's first img row
// (the centre of strip row i is img row i + 3).  A FAST-9 corner at threshold t has an arc of 9 ring pixels with |v - r| > t; an arc of 9
// holds one pixel of every opposite pair (k, k + 8), so |v - r| > t must hold on at least one pixel of each of the four pairs of the even
// ring positions: (0, +-3), (+-3, 0), (2, 2)/(-2, -2), (2, -2)/(-2, 2).  Byte-SIMD: 4 pixels per VABSDIFF4, the compare is the carry into
// bit 7 of d + (127 - t) (or d >= 128); flags of strip rows 0-6 go to mA, rows 7-12 to mB (byte = pixel, bit = row).
// (On the bench images: 12.5 % of the pixels survive, 4.2 % are corners; the former test on the pairs (0, 8) and (4, 12) alone left 23.5 %.)
// (Moving the byte funnel shifts and the flag gather to the FMA pipe — mul.hi + mad instead of PRMT, mad.hi instead of SHF + LOP3 — was
//  measured slower, 0.73 vs 0.705 ms: the ALU pipe drops from 65 % to 56 % but the kernel is bound by issue slots, not by that pipe.)
template <int IP, bool FULL>
__device__ __forceinline__ void fast_quick8(uint32_t col, int rs, uint32_t cadd, uint32_t &mA, uint32_t &mB) {
    uint32_t cu[kFastStrip + 6], dm[kFastStrip + 6], dp[kFastStrip + 6], hl[kFastStrip + 6], hr[kFastStrip + 6];
    mA = 0; mB = 0;
#pragma unroll
    for (int j = 0; j < kFastStrip + 6; j++) {
        if (FULL || j < 7 || j < rs + 6) {               // (warp-uniform; FULL: rs == kFastStrip)
            const uint32_t rp = col + j * IP;
            cu[j] = lds32_ro(rp);
            if (j >= 1 && j <= kFastStrip + 4) {         // rows that serve as y - 2, y, y + 2 of some centre row
                const uint32_t pv = lds32_ro(rp - 4), nx = lds32_ro(rp + 4);
                dm[j] = __byte_perm(pv, cu[j], 0x5432);  // pixels x - 2
                dp[j] = __byte_perm(cu[j], nx, 0x5432);  // pixels x + 2
                if (j >= 3 && j <= kFastStrip + 2) { hl[j] = __byte_perm(pv, cu[j], 0x4321); hr[j] = __byte_perm(cu[j], nx, 0x6543); }
            }
            if (j >= 6) {
                const int i = j - 6;                     // centre row of the strip, img row i + 3
                const uint32_t v = cu[i + 3];
                const uint32_t d0 = __vabsdiffu4(v, cu[i]), d8 = __vabsdiffu4(v, cu[i + 6]);
                const uint32_t d4 = __vabsdiffu4(v, hr[i + 3]), d12 = __vabsdiffu4(v, hl[i + 3]);
                const uint32_t d2 = __vabsdiffu4(v, dp[i + 5]), d10 = __vabsdiffu4(v, dm[i + 1]);
                const uint32_t d6 = __vabsdiffu4(v, dp[i + 1]), d14 = __vabsdiffu4(v, dm[i + 5]);
                const uint32_t acc = (d0 | d8 | (d0 + cadd) | (d8 + cadd)) & (d4 | d12 | (d4 + cadd) | (d12 + cadd)) &
                                     (d2 | d10 | (d2 + cadd) | (d10 + cadd)) & (d6 | d14 | (d6 + cadd) | (d14 + cadd));
                if (i < kFastHalf) mA |= (acc >> (7 - i)) & (0x01010101u << i);
                else mB |= (acc >> (7 - (i - kFastHalf))) & (0x01010101u << (i - kFastHalf));
            }
        }
    }
}

// Scores the centre pixels at img offsets a0 / a1 (queue entries) with packed u16x2 arithmetic, stores S where S >= t and
// appends those pixels to the warp's corner list at shared address cl (ncl = warp-uniform fill level).  img / scb: shared addresses of
// the image tile and of the score of img offset 0.
template <int IP>
__device__ __forceinline__ void fast_score_pair(uint32_t img, uint32_t scb, int a0, int a1, bool ok0, bool ok1, int t,
                                                uint32_t cl, int &ncl, unsigned lt_mask) {
    const uint32_t p0 = img + a0, p1 = img + a1;
    constexpr int o[16] = {3 * IP, 3 * IP + 1, 2 * IP + 2, IP + 3, 3, -IP + 3, -2 * IP + 2, -3 * IP + 1,
                           -3 * IP, -3 * IP - 1, -2 * IP - 2, -IP - 3, -3, IP - 3, 2 * IP - 2, 3 * IP - 1};
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // pack on the FMA pipe (IMAD), not with PRMT / LOP3: the kernel is bound by the ALU pipe (VIMNMX3, VABSDIFF4, LOP3)
        asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r[k]) : "r"(lds8_ro(p1 + o[k])), "r"(lds8_ro(p0 + o[k])));
    }
    uint32_t mx3[16], mn3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mx3[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mn3[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    uint32_t A = 0xFFFFFFFFu, B = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t x0 = __vimax3_u16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
        const uint32_t x1 = __vimax3_u16x2(mx3[k + 1], mx3[(k + 4) & 15], mx3[(k + 7) & 15]);
        A = __vimin3_u16x2(A, x0, x1);
        const uint32_t y0 = __vimin3_u16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
        const uint32_t y1 = __vimin3_u16x2(mn3[k + 1], mn3[(k + 4) & 15], mn3[(k + 7) & 15]);
        B = __vimax3_u16x2(B, y0, y1);
    }
    // S = max(v - A, B - v) - 1 for both halves at once: v - A and B - v never wrap when they matter (A <= v or B >= v for a corner),
    // a wrapped half is > 255 as u16 only if it was negative, in which case the other term or the threshold test discards it
    const int v0 = (int)lds8_ro(p0), v1 = (int)lds8_ro(p1);
    const int s0 = max(v0 - (int)(A & 0xFFFFu), (int)(B & 0xFFFFu) - v0) - 1;
    const int s1 = max(v1 - (int)(A >> 16), (int)(B >> 16) - v1) - 1;
    const bool c0 = ok0 && s0 >= t, c1 = ok1 && s1 >= t;
    if (c0) sts8(scb + a0, (uint32_t)s0);
    if (c1) sts8(scb + a1, (uint32_t)s1);
    const unsigned b0 = __ballot_sync(0xFFFFFFFFu, c0), b1 = __ballot_sync(0xFFFFFFFFu, c1);
    const int n0 = __popc(b0);
    {   // (positions for every lane, predicated stores: as two divergent blocks this tail cost 10 more instructions per call)
        const int p0 = ncl + __popc(b0 & lt_mask), p1 = ncl + n0 + __popc(b1 & lt_mask);
        if (c0 && p0 < kFastCorners) sts16(cl + 2 * p0, (uint32_t)a0);
        if (c1 && p1 < kFastCorners) sts16(cl + 2 * p1, (uint32_t)a1);
    }
    ncl += n0 + __popc(b1);
}

// strict 3x3 maximum of the score at (tile row, img column) inside its cell (columns [cx0, cx0 + wc), wdt of them tested) -> keep bitmap
// of the cell (bit = row * wdt + column inside the cell); sc1 / kb: shared addresses of the score of tile row 0, column 0 and of the bitmap
template <int IP>
__device__ __forceinline__ bool fast_nms_one(uint32_t sc1, uint32_t kb, int row, int col, int cx0, int wc, int wdt) {
    const uint32_t n = sc1 + row * IP + col;
    const int xl = col - cx0;
    // all nine scores are loaded before the first compare (independent loads instead of a short-circuit chain)
    const int s = lds8(n), u = lds8(n - IP), d = lds8(n + IP), l = lds8(n - 1), ul = lds8(n - IP - 1), dl = lds8(n + IP - 1), r = lds8(n + 1),
              ur = lds8(n - IP + 1), dr = lds8(n + IP + 1);
    const int ml = xl > 0 ? max(l, max(ul, dl)) : -1, mr = xl < wc - 1 ? max(r, max(ur, dr)) : -1;
    const bool keep = s > max(max(u, d), max(ml, mr));
    const int bit = row * wdt + xl;
    const uint32_t ka = kb + 4 * (bit >> 5), kv = 1u << (bit & 31);
    if (keep) reds_or32(ka, kv);
    return keep;
}

// pyramid levels as TMA tensors of 16-bit elements for the FAST tiles: dims (pitch / 2, plane rows, images), box = (fast_ip / 2, fast_ymax + 6, 1)
struct FastMaps { CUtensorMap m[kMaxLevels]; int use_tma; };

// optional per-launch statistics (profiling mode only): tested pixels, quick-test survivors, corners, kept after NMS, retried cells, tiles
struct FastStats { unsigned long long tested, survivors, corners, kept, retry_cells, tiles; };

struct FastArgs { int ymax, total_cells, img0, kb_words; long long cand_per_image; };

// ceil(2^32 / d) for 2 <= d < 128: x / d == umulhi(x, magic) for x * d < 2^32 (d == 1 is handled by the caller)
struct DivMagicTable {
    uint32_t v[128];
    constexpr DivMagicTable() : v() { for (int d = 1; d < 128; d++) v[d] = 0xFFFFFFFFu / (uint32_t)d + 1u; }
};
__constant__ DivMagicTable kDivMagic = DivMagicTable();

template <int IP>
__global__ void __launch_bounds__(kFastThreads, 8) k_fast(const uint8_t *__restrict__ pyr, const FastTile *__restrict__ tiles, const FastArgs A,
                                                       const orbx_image_params *__restrict__ params,
                                                       uint32_t *__restrict__ cells, int *__restrict__ cell_cnt,
                                                       const __grid_constant__ FastMaps maps, FastStats *__restrict__ stats) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int b = blockIdx.y + A.img0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // tile descriptor: the same for every thread (broadcast loads)
    const FastTile *tile = tiles + blockIdx.x;
    const int4 *tp = reinterpret_cast<const int4 *>(tile);
    const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2), t3 = __ldg(tp + 3);
    const int lvl = t0.x, ncell = t0.y, wc = t0.z, cell_cap = t0.w;
    const int tx0 = t1.x, ty0 = t1.y, X = t1.z, Y = t1.w;
    const int pc0 = t2.x, prow0 = t2.y, c0 = t2.z, cell0 = t2.w;
    const int nstr = t3.x, rs = t3.y;
    if (X <= 0 || Y <= 0) {                              // nothing to test (degenerate last cell row / column)
        if (tid < ncell) cell_cnt[(long long)b * A.total_cells + cell0 + tid] = 0;
        return;
    }
    const int img_bytes = IP * (A.ymax + 6), zero_bytes = fast_zero_bytes(IP, A.ymax);
    // ---- stage the tile: plane rows prow0 .. prow0 + Y + 5, 16-byte columns from pc0
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned long long s_stats[4];           // (profiling mode) survivors, corners, kept, retried cells of this tile
    if (tid < 4) s_stats[tid] = 0;
    if (maps.use_tma) {
        // ONE TMA tensor copy stages the tile (box = fast_ip x (fast_ymax + 6) from plane column pc0, row prow0; what lies
        // beyond the columns / rows this tile needs is harmless), in flight while the CTA clears the score map
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)img_bytes);
            tma_load_3d(smem, &maps.m[lvl], pc0 >> 1, prow0, b, &s_bar);   // x in 16-bit elements
        }
    } else {
        // 16-byte cp.async (LDGSTS) copies.  (One TMA bulk copy per tile ROW was measured slower than this: ~190 small requests
        // per tile time saturate the SM's TMA unit; one tensor copy per TILE is faster than both.)
        const int4 t4 = __ldg(tp + 4), t5 = __ldg(tp + 5);
        const long long plane_off = ((long long)(unsigned)t4.x) | ((long long)t4.y << 32), plane_bytes = ((long long)(unsigned)t4.z) | ((long long)t4.w << 32);
        const int pitch = t5.x;
        const int nvec = (c0 + X + 3 + 15) >> 4;        // uint4 per row
        if (lane < nvec) {
            const uint8_t *src = pyr + plane_off + (long long)b * plane_bytes + (long long)prow0 * pitch + pc0 + 16 * lane;
            const uint32_t dst = smem_u32(smem) + 16 * lane;
#pragma unroll 1
            for (int r = wid; r < Y + 6; r += kFastWarps)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * IP), "l"(src + (long long)r * pitch) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // shared addresses, opaque to the compiler from here on: image tile, score of img offset 0 / of tile row 0, this warp's region
    // [q: kFastQueue u16][cl: kFastCorners u16][kb: kb_words u32]
    const uint32_t sc = smem_u32(smem) + img_bytes;
    {
#pragma unroll 2
        for (int i = tid; i < (zero_bytes >> 4); i += kFastThreads) sts128_zero(sc + 16 * i);   // score map
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (maps.use_tma && tid < 32) mbar_wait(&s_bar, 0);
    __syncthreads();
    const uint32_t img = sa_keep(smem_u32(smem));
    const uint32_t scb = sa_keep(sc - 2 * IP);          // score of the pixel at img offset a: scb + a (img row = tile row + 3)
    const uint32_t q = sa_keep(sc + zero_bytes + wid * fast_warp_bytes(A.kb_words));
    const uint32_t cl = q + 2 * kFastQueue, kb = cl + 2 * kFastCorners;
    const uint32_t sc1 = scb + 3 * IP;                  // score of tile row 0, column 0
    // ---- from here on every warp works alone on its own cells: quick test -> queue -> scores -> NMS -> (retry) -> emission.
    //      Cells only share the read-only image tile; scores and keep bits of a cell are written and read by its warp only.
    const int t_ini = min(max(__ldg(&params[b].ini_th_fast), 0), 255), t_min = min(max(__ldg(&params[b].min_th_fast), 0), 255);
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll 1
    for (int k = wid; k < ncell; k += kFastWarps) {
        int *cnt_out = cell_cnt + (long long)b * A.total_cells + cell0 + k;
        const int cx0 = c0 + k * wc, cx1 = min(cx0 + wc, c0 + X);   // smem columns of the cell's tested pixels
        if (cx1 <= cx0) { if (lane == 0) *cnt_out = 0; continue; }
        const int wdt = cx1 - cx0;
        const int w0 = cx0 >> 2, w1 = (cx1 - 1) >> 2;   // words of a row that hold tested pixels of this cell
        const int nw = w1 - w0 + 1;
        const float inv_nw = __fdividef(1.0f, (float)nw);   // (2 ulp are far inside small_div's margin; the IEEE division is 20 instructions)
        const int nflat = nstr * nw;                    // work items: (strip of rs rows) x (word column)
        const int nkb = (Y * wdt + 31) >> 5;            // words of the cell's keep bitmap
        bool any_kept = false;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            const int t = pass == 0 ? t_ini : t_min;
            // per byte: x + cadd has bit 7 set iff x > t for x < 128 (x >= 128 > t is caught by OR-ing x itself, t <= 126).  A carry out
            // of a byte can only turn |d| == t of the next byte into a survivor: the quick test is a necessary condition, scoring is exact.
            const uint32_t cadd = (uint32_t)(0x80 - (min(t, 126) + 1)) * 0x01010101u;
            if (pass == 0) {                            // (a repeated cell kept nothing: still clear)
#pragma unroll 1
                for (int i = lane; i < ((nkb + 1) & ~1); i += 32) sts32(kb + 4 * i, 0u);
            }
            // ---- 1 + 2: quick reject, survivor queue, pair scoring.  The last trip of the loop (f0 >= nflat) only drains the queue.
            int qn = 0, ncl = 0;                        // fill levels of this warp's queue and corner list (warp-uniform)
#pragma unroll 1
            for (int f0 = 0;; f0 += 32) {
                const bool drain = f0 >= nflat;
                const int f = f0 + lane;
                uint32_t mA = 0, mB = 0;
                int abase = 0;
                if (!drain && f < nflat) {
                    const int strip = small_div(f, inv_nw), wi = w0 + (f - strip * nw);
                    int own0, own1;                     // rows this strip owns; it marches rs rows from r0 <= own0 (never past row Y + 2)
                    if (nstr <= kFastMaxStrips) { own0 = __ldg(&tile->sr[strip]); own1 = __ldg(&tile->sr[strip + 1]); }
                    else { own0 = strip * rs; own1 = min(own0 + rs, Y); }
                    const int r0 = min(own0, Y - rs);
                    // pixels of this word that belong to the cell
                    const int first = max(cx0 - 4 * wi, 0), last = min(cx1 - 1 - 4 * wi, 3);
                    const uint32_t cm = (0xFFFFFFFFu << (8 * first)) & (0xFFFFFFFFu >> (8 * (3 - last)));
                    abase = (r0 + 3) * IP + 4 * wi;
                    if (rs == kFastStrip) fast_quick8<IP, true>(img + r0 * IP + 4 * wi, rs, cadd, mA, mB);
                    else fast_quick8<IP, false>(img + r0 * IP + 4 * wi, rs, cadd, mA, mB);
                    const uint32_t rows = ((1u << min(rs, own1 - r0)) - 1u) & ~((1u << (own0 - r0)) - 1u);   // bit i = strip row i is owned
                    mA &= cm & (0x01010101u * (rows & ((1u << kFastHalf) - 1u)));
                    mB &= cm & (0x01010101u * (rows >> kFastHalf));
                }
                if (stats) { const int c = __reduce_add_sync(0xFFFFFFFFu, __popc(mA) + __popc(mB)); if (lane == 0) reds_add64(smem_u32(&s_stats[0]), (unsigned long long)c); }
                // append the survivors of both flag registers to the warp's queue (entries are img byte offsets, the order inside the queue
                // is irrelevant): one warp scan, then every lane lists its flags, one of each register per trip.  More survivors than the
                // queue holds (rare: > 448 of 1792 pixels) are appended row by row.
                {
                    const int nsurv = __reduce_add_sync(0xFFFFFFFFu, __popc(mA) + __popc(mB));
                    const int nsub = nsurv == 0 ? 0 : nsurv <= kFastQueue - 64 ? 1 : kFastHalf;
                    const int need = drain ? 1 : 64;
#pragma unroll 1
                    for (int sub = 0; sub < max(nsub, 1); sub++) {
                        if (nsub) {
                            uint32_t fa = nsub == 1 ? mA : mA & (0x01010101u << sub), fb = nsub == 1 ? mB : mB & (0x01010101u << sub);
                            const int ca = __popc(fa), cnt = ca + __popc(fb);
                            int inc = cnt;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += u; }
                            uint32_t qa = q + 2 * (qn + inc - cnt), qb = qa + 2 * ca;
                            qn += __shfl_sync(0xFFFFFFFFu, inc, 31);
                            const int ab2 = abase + kFastHalf * IP;
                            // branch-free listing: every lane walks its 28 + 24 flag positions once (flag test, predicated store, predicated
                            // pointer bump, entry value: 4 instructions per position, 208 per append).  The former next-set-bit loop ran as
                            // many trips as the fullest lane had flags — 12 on average with 7 of 32 lanes active, 450 instructions per
                            // append (round 2: k_fast 0.729 -> 0.661 ms per 320 images).
#pragma unroll
                            for (int px = 0; px < 4; px++) {
#pragma unroll
                                for (int i = 0; i < kFastHalf; i++) {
                                    if (fa & (1u << (8 * px + i))) { sts16(qa, (uint32_t)(abase + i * IP + px)); qa += 2; }
                                    if (kFastHalf + i < kFastStrip && (fb & (1u << (8 * px + i)))) { sts16(qb, (uint32_t)(ab2 + i * IP + px)); qb += 2; }
                                }
                            }
                            __syncwarp();
                        }
                        while (qn >= need) {
                            const int n = min(qn, 64);
                            qn -= n;
                            const bool ok0 = lane < n, ok1 = lane + 32 < n;
                            const int a0 = (int)lds16(q + 2 * (qn + (ok0 ? lane : 0))), a1 = (int)lds16(q + 2 * (qn + (ok1 ? lane + 32 : 0)));
                            fast_score_pair<IP>(img, scb, a0, a1, ok0, ok1, t, cl, ncl, lt_mask);
                        }
                        __syncwarp();
                    }
                }
                if (drain) break;
            }
            // ---- 3: strict 3x3 maximum inside the cell, over the corner list (or a sweep of the cell's score map if it overflowed)
            bool kept = false;
            if (ncl <= kFastCorners) {
#pragma unroll 1
                for (int i = lane; i < ncl; i += 32) {
                    const int a = (int)lds16(cl + 2 * i);
                    const int ri = a / IP;
                    kept |= fast_nms_one<IP>(sc1, kb, ri - 3, a - ri * IP, cx0, wc, wdt);
                }
            } else {
                const int items = Y * nw;
#pragma unroll 1
                for (int it = lane; it < items; it += 32) {
                    const int row = small_div(it, inv_nw), wi = w0 + (it - row * nw);
                    const uint32_t wv = lds32(sc1 + row * IP + 4 * wi);
                    if (wv == 0) continue;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
                        if (((wv >> (8 * kk)) & 0xFF) && 4 * wi + kk >= cx0 && 4 * wi + kk < cx1)
                            kept |= fast_nms_one<IP>(sc1, kb, row, 4 * wi + kk, cx0, wc, wdt);
                }
            }
            any_kept = __any_sync(0xFFFFFFFFu, kept);
            __syncwarp();
            if (stats && lane == 0) reds_add64(smem_u32(&s_stats[1]), (unsigned long long)ncl);
            // ---- 4: a cell that found nothing is repeated with minThFAST (src/ORBextractor.cc:760-782)
            if (any_kept || t_min >= t_ini) break;
            if (pass == 0 && stats && lane == 0) reds_add64(smem_u32(&s_stats[3]), 1ull);
        }
        // ---- 5: emission in cv::FAST order (row-major inside the cell) = ascending bit index of the cell's keep bitmap: a lane takes
        //      64 bits, one warp scan per 2048 bits (a 37 x 37 cell has 1369)
        int total = 0;
        if (any_kept) {
            const long long cand0 = ((long long)(unsigned)t3.z) | ((long long)t3.w << 32);
            uint32_t *out = cells + (long long)b * A.cand_per_image + cand0 + (long long)k * cell_cap;
            // row of bit x = x / wdt as a multiply-high (exact for x * wdt < 2^32; x < 2^16 here)
            const uint32_t magic = wdt < 128 ? kDivMagic.v[wdt] : 0xFFFFFFFFu / (uint32_t)wdt + 1u;   // (table: the division is 25 instructions)
            // border-relative coordinates (src/ORBextractor.cc:784-791): level coordinate - 16
            const uint32_t base = (uint32_t)(tx0 + k * wc - kFastBorder) | ((uint32_t)(ty0 - kFastBorder) << 12);
            const uint32_t scc = sc1 + cx0;
#pragma unroll 1
            for (int w2 = 2 * lane; w2 - 2 * lane < nkb; w2 += 64) {     // (warp-uniform trip count; one trip for cells up to 2048 pixels)
                uint32_t v0 = w2 < nkb ? lds32(kb + 4 * w2) : 0u, v1 = w2 + 1 < nkb ? lds32(kb + 4 * w2 + 4) : 0u;
                const int n = __popc(v0) + __popc(v1);
                int inc = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += u; }
                const int nchunk = __shfl_sync(0xFFFFFFFFu, inc, 31);
                // the lanes list the set bits of their 64 in order (a short serial loop; the corner list is free again), then the list is
                // emitted one keypoint per lane; what does not fit the list (> kFastCorners kept pixels in 2048) is written directly
                int li = inc - n;
#pragma unroll 1
                while (v0 | v1) {
                    const bool lo = v0 != 0;
                    const uint32_t vv = lo ? v0 : v1;
                    const int bit = 32 * w2 + (lo ? 0 : 32) + __ffs(vv) - 1;
                    if (lo) v0 &= v0 - 1; else v1 &= v1 - 1;
                    if (li < kFastCorners) sts16(cl + 2 * li, (uint32_t)bit);
                    else if (total + li < cell_cap) {
                        const int row = wdt == 1 ? bit : (int)__umulhi((uint32_t)bit, magic), xc = bit - row * wdt;   // (2^32 / 1 has no 32-bit magic)
                        out[total + li] = (base + (uint32_t)xc + ((uint32_t)row << 12)) | (lds8(scc + row * IP + xc) << 24);
                    }
                    li++;
                }
                __syncwarp();
                const int nlist = min(min(nchunk, kFastCorners), cell_cap - total);
#pragma unroll 1
                for (int j = lane; j < nlist; j += 32) {
                    const int bit = (int)lds16(cl + 2 * j);
                    const int row = wdt == 1 ? bit : (int)__umulhi((uint32_t)bit, magic), xc = bit - row * wdt;   // (2^32 / 1 has no 32-bit magic)
                    out[total + j] = (base + (uint32_t)xc + ((uint32_t)row << 12)) | (lds8(scc + row * IP + xc) << 24);
                }
                __syncwarp();
                total += nchunk;
            }
        }
        if (lane == 0) *cnt_out = min(total, cell_cap);
        if (stats && lane == 0) reds_add64(smem_u32(&s_stats[2]), (unsigned long long)total);
    }
    if (stats) {
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&stats->survivors, s_stats[0]); atomicAdd(&stats->corners, s_stats[1]); atomicAdd(&stats->kept, s_stats[2]);
            atomicAdd(&stats->retry_cells, s_stats[3]);
            atomicAdd(&stats->tested, (unsigned long long)X * Y); atomicAdd(&stats->tiles, 1ull);
        }
    }
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
    __device__ __forceinline__ void hist_add(int *cnt, int key, bool valid, int nkeys) {
        if (nkeys <= 32) {
            // thousands of candidates share these few counters: one atomic per (warp, key).  Lane k counts key k from bit-sliced
            // ballots (one per key BIT plus the valid mask: at most 6 ballots instead of one ballot per key)
            unsigned m = __ballot_sync(0xFFFFFFFFu, valid);
            const int lane = tid & 31;
            for (int j = 0; (1 << j) < nkeys; j++) {
                const unsigned bj = __ballot_sync(0xFFFFFFFFu, (key >> j) & 1);
                m &= ((lane >> j) & 1) ? bj : ~bj;
            }
            if (lane < nkeys && m) atomicAdd(&cnt[lane], __popc(m));
        } else if (valid) {
            atomicAdd(&cnt[key], 1);
        }
    }
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

__device__ __forceinline__ void octree_body(const Geom &G, const int *__restrict__ nfeat_lvl, const uint32_t *__restrict__ cells,
                                            const int *__restrict__ cell_cnt, int *__restrict__ cell_off,
                                            uint32_t *__restrict__ lin, int *__restrict__ lin_n,
                                            unsigned short *__restrict__ node, int *__restrict__ sel_tmp,
                                            uint32_t *__restrict__ lvl_kp, int *__restrict__ lvl_n, int list_cap, int img0) {
    extern __shared__ __align__(16) unsigned char oct_smem[];
    __shared__ int s_total;
    // grid = (images, levels): CTAs are dispatched level 0 first — the levels with the most candidates and the longest
    // sequential phases start first, the short ones fill the tail
    const int lvl = blockIdx.y, b = blockIdx.x + img0;
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

// throughput shape: 256 threads, 6 CTAs per SM (large batches keep every SM busy with independent (image, level) pairs)
__global__ void __launch_bounds__(256, 6) k_octree(Geom G, const int *__restrict__ nfeat_lvl, const uint32_t *__restrict__ cells,
                                                const int *__restrict__ cell_cnt, int *__restrict__ cell_off,
                                                uint32_t *__restrict__ lin, int *__restrict__ lin_n,
                                                unsigned short *__restrict__ node, int *__restrict__ sel_tmp,
                                                uint32_t *__restrict__ lvl_kp, int *__restrict__ lvl_n, int list_cap, int img0) {
    octree_body(G, nfeat_lvl, cells, cell_cnt, cell_off, lin, lin_n, node, sel_tmp, lvl_kp, lvl_n, list_cap, img0);
}
// latency shape: 1024 threads per (image, level) for batches that leave most SMs idle (one rig frame): the candidate sweeps of
// the long level-0 chain get four times the threads
__global__ void __launch_bounds__(1024, 1) k_octree_wide(Geom G, const int *__restrict__ nfeat_lvl, const uint32_t *__restrict__ cells,
                                                      const int *__restrict__ cell_cnt, int *__restrict__ cell_off,
                                                      uint32_t *__restrict__ lin, int *__restrict__ lin_n,
                                                      unsigned short *__restrict__ node, int *__restrict__ sel_tmp,
                                                      uint32_t *__restrict__ lvl_kp, int *__restrict__ lvl_n, int list_cap, int img0) {
    octree_body(G, nfeat_lvl, cells, cell_cnt, cell_off, lin, lin_n, node, sel_tmp, lvl_kp, lvl_n, list_cap, img0);
}

// ------------------------------------------------------------------------------------------------------
// K7a  output slots (src/ORBextractor.cc:1024-1068): levels ascending, keypoints in list order; a keypoint
// whose scaled x lies in [lap0, lap1] takes slot stereoIndex-- (from n-1 down), the others monoIndex++.
// One block per image.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slots(Geom G, const orbx_image_params *__restrict__ params,
                                               const uint32_t *__restrict__ lvl_kp, const int *__restrict__ lvl_n,
                                               int *__restrict__ slot, int *__restrict__ n_kp, int *__restrict__ n_mono, int img0) {
    __shared__ int s_warp[8];
    const int b = blockIdx.x + img0;
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
// TMA tensor maps of the pyramid / blurred-pyramid levels with per-keypoint boxes: 32 x 31 (orientation patch) and 48 x 37
// (descriptor reach +-18) bytes.  The x coordinate of a box is rounded down to 16 bytes (a misaligned inner coordinate raised an
// illegal-instruction fault on B200), the remainder is added to the column index inside the staged patch.
struct PatchMaps { CUtensorMap pyr[kMaxLevels], blur[kMaxLevels]; int use_tma; };
constexpr int kOriW = 48, kOriH = 31, kDescW = 64, kDescH = 37;   // widths include up to 15 columns of 16-byte alignment slack
constexpr int kOriBytes = kOriW * kOriH, kDescBytes = kDescW * kDescH;
constexpr int kPatchStride = ((kOriBytes + 127) & ~127) + ((kDescBytes + 127) & ~127);      // per warp, 128-byte aligned parts

// IC_Angle, one patch row per lane: the lane holds its whole staged 48-byte row as twelve words (three 128-bit loads: with a row pitch
// of 48 bytes the eight lanes of a quarter warp hit eight different 16-byte bank groups, so the loads are conflict-free, while the nine
// 32-bit loads per lane this replaces ran into 4-way conflicts — 36 instead of 12 shared-memory wavefronts per keypoint, in a kernel that
// is bound by the L1 / shared-memory pipe).  The 31 pixels u = -15..15 start at byte c0 of the row (1 <= c0 <= 16): words kS = c0 >> 2
// onwards, funnel-shifted by sh = 8 * (c0 & 3); msk = the disc as byte masks; m10 += sum u I (IDP.4A, signed weights), rowsum += sum I.
template <int kS>
__device__ __forceinline__ void ic_row_sums(const uint32_t (&w)[12], int sh, const uint32_t (&msk)[8], int &m10, int &rowsum) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t a = __funnelshift_r(w[kS + q], w[kS + q + 1 < 12 ? kS + q + 1 : 11], sh) & msk[q];   // (kS == 4: sh == 0, the high word is unused)
        const int u0 = 4 * q - kHalfPatch;                       // weights u0 .. u0 + 3 as signed bytes
        const uint32_t wu = (uint32_t)(u0 & 0xFF) | ((uint32_t)((u0 + 1) & 0xFF) << 8) | ((uint32_t)((u0 + 2) & 0xFF) << 16) |
                            ((uint32_t)((u0 + 3) & 0xFF) << 24);
        asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(a), "r"(wu));
        rowsum = (int)__dp4a(a, 0x01010101u, (uint32_t)rowsum);
    }
}
__device__ __forceinline__ void ic_row(uint32_t row_addr, int c0, const uint32_t (&msk)[8], int &m10, int &rowsum) {
    uint32_t w[12];
#pragma unroll
    for (int k = 0; k < 3; k++)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * k]), "=r"(w[4 * k + 1]), "=r"(w[4 * k + 2]), "=r"(w[4 * k + 3]) : "r"(row_addr + 16 * k) : "memory");
    const int sh = 8 * (c0 & 3);
    switch (c0 >> 2) {                                           // warp-uniform
    case 0: ic_row_sums<0>(w, sh, msk, m10, rowsum); break;
    case 1: ic_row_sums<1>(w, sh, msk, m10, rowsum); break;
    case 2: ic_row_sums<2>(w, sh, msk, m10, rowsum); break;
    case 3: ic_row_sums<3>(w, sh, msk, m10, rowsum); break;
    default: ic_row_sums<4>(w, sh, msk, m10, rowsum); break;
    }
}

// grid.x enumerates CTAs of 8 keypoint slots; a CTA belongs to one level: block_info[blockIdx.x] = level | first slot << 4
template <bool kFma, bool kTma>
__global__ void __launch_bounds__(256) k_orient_desc(const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur, Geom G,
                                                     const signed char *__restrict__ pattern, const uint32_t *__restrict__ lvl_kp,
                                                     const int *__restrict__ lvl_n, const int *__restrict__ slot,
                                                     float *__restrict__ lvl_angle, orbx_kp *__restrict__ out_kp,
                                                     uint8_t *__restrict__ out_desc, const PatchMaps *__restrict__ maps,
                                                     const int *__restrict__ block_info, int img0) {
    __shared__ __align__(128) uint8_t s_patch[kTma ? 8 * kPatchStride : 128];
    __shared__ __align__(8) uint64_t s_bar[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y + img0;
    const int info = __ldg(block_info + blockIdx.x);
    const int lvl = info & 15;
    const LevelGeom &L = G.lv[lvl];
    const int i = (info >> 4) + wid;
    const int warp = L.kp_off + i;                              // slot of this keypoint in the per-image level arena
    if (i >= lvl_n[b * G.nlevels + lvl]) return;
    const uint32_t kp = lvl_kp[(long long)b * G.kp_per_image + warp];
    const int x = cand_x(kp), y = cand_y(kp);
    uint8_t *op = s_patch + wid * kPatchStride;                 // rows y-15..y+15, columns from (x-16) & ~15 of the level
    uint8_t *dp = op + ((kOriBytes + 127) & ~127);              // rows y-18..y+18, columns from (x-24) & ~15 of the blurred level
    constexpr bool tma = kTma;               // (the maps live in global memory: as kernel parameters they would push the
                                             //  parameter block beyond 4 KB)
    if (tma) {   // the warp fetches its two patches with two TMA tensor copies and waits on its own mbarrier
        if (lane == 0) {
            mbar_init(&s_bar[wid], 1);
            mbar_arrive_expect_tx(&s_bar[wid], (uint32_t)(kOriBytes + kDescBytes));
            tma_load_3d(op, &maps->pyr[lvl], (kXO + x - 16) & ~15, kYO + y - kHalfPatch, b, &s_bar[wid]);
            tma_load_3d(dp, &maps->blur[lvl], (kXO + x - 24) & ~15, kYO + y - 18, b, &s_bar[wid]);
        }
        __syncwarp();
        mbar_wait(&s_bar[wid], 0);
    }
    // ---- orientation (IC_Angle): m10 = sum u I, m01 = sum v I over the disc |u| <= umax[|v|]
    int m10 = 0, m01 = 0;
    if (tma) {
        // lane = row v of the staged patch: its 31 bytes u = -15..15 are 8 funnel-shifted words; the disc is a byte mask, the two
        // sums are IDP.4A with the weights u (signed) and 1
        if (lane < 31) {
            const int v = lane - kHalfPatch, av = v < 0 ? -v : v;
            const int m = (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15);                      // umax[|v|], src/ORBextractor.cc:399-413
            const uint32_t bits = ((2u << (2 * m)) - 1u) << (kHalfPatch - m);                    // bytes 15 - m .. 15 + m of the row
            const int c0 = ((kXO + x - 16) & 15) + 1;                                            // byte of column x - 15 in the staged row
            // (nine 32-bit loads per lane: with the 48-byte row pitch they run into 4-way bank conflicts; three conflict-free 128-bit
            //  loads of the whole row + a warp-uniform switch over the first word — ic_row below, used by the batched kernel — cost 8
            //  more registers here (6 instead of 7 CTAs per SM) and were measured slower in this kernel: 0.343 vs 0.327 ms)
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(op + lane * kOriW + (c0 & ~3));
            const int sh = 8 * (c0 & 3);
            uint32_t w[9];
#pragma unroll
            for (int j = 0; j < 9; j++) w[j] = rw[j];
            int rowsum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t px4 = __funnelshift_r(w[j], w[j + 1], sh);
                const uint32_t nib = (bits >> (4 * j)) & 0xFu;
                const uint32_t msk = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
                const uint32_t a = px4 & msk;
                const int u0 = 4 * j - kHalfPatch;                                               // weights u0 .. u0 + 3 as signed bytes
                const uint32_t wu = (uint32_t)(u0 & 0xFF) | ((uint32_t)((u0 + 1) & 0xFF) << 8) | ((uint32_t)((u0 + 2) & 0xFF) << 16) |
                                    ((uint32_t)((u0 + 3) & 0xFF) << 24);
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(a), "r"(wu));
                rowsum = (int)__dp4a(a, 0x01010101u, (uint32_t)rowsum);
            }
            m01 = v * rowsum;
        }
    } else if (lane < 31) {
        // lane = column u of the patch; the lane's column sum gives m10, the row-weighted sum m01
        const int u = lane - kHalfPatch;
        const int au = u < 0 ? -u : u;
        constexpr int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};   // src/ORBextractor.cc:399-413
        int colsum = 0;
        const uint8_t *p = plane_ptr(pyr, L, b) + (long long)(kYO + y - kHalfPatch) * L.pitch + kXO + x + u;
#pragma unroll
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) {   // fully unrolled: 31 independent loads in flight
            if (au <= umax[v < 0 ? -v : v]) {
                const int val = __ldg(p);
                colsum += val; m01 += v * val;
            }
            p += L.pitch;
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
    // (Evaluating atan2 + the glibc-exact sincos once per CTA for its 8 keypoints — lane j of warp 0 takes keypoint j, two CTA barriers —
    //  was measured slower, 0.344 vs 0.326 ms: the barriers cost more than the ~100 saved warp-instructions per keypoint.)
    const float angle = orbx_fast_atan2((float)m01, (float)m10);
    const float rad = __fmul_rn(angle, 0.0174532924f);   // (float)(CV_PI/180.f)
    float sn, cs;
    orbx_sincosf(rad, &sn, &cs);
    // ---- descriptor: lane = output byte; its 16 pattern points are 32 signed bytes = two 128-bit loads (a float table — 8 float4
    //      per lane — was measured 2x slower: 128 B per lane of strided L1 traffic)
    const uint4 pa = __ldg(reinterpret_cast<const uint4 *>(pattern) + 2 * lane), pb = __ldg(reinterpret_cast<const uint4 *>(pattern) + 2 * lane + 1);
    const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};   // word k = (x0, y0, x1, y1) of test k
    const uint8_t *cb = tma ? dp + 18 * kDescW + 24 + ((kXO + x - 24) & 15) : plane_ptr(blur, L, b) + (long long)(kYO + y) * L.pitch + kXO + x;
    const int pitch = tma ? kDescW : L.pitch;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int tv[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            // (int8 -> float and cvRound through the 1.5 * 2^23 magic constant — PRMT + IADD + FADD instead of I2F / F2I — was measured
            //  slower: the kernel is bound by issue slots, not by the conversion pipe)
            const float px = (float)(signed char)(pw[k] >> (16 * e)), py = (float)(signed char)(pw[k] >> (16 * e + 8));
            float fy, fx;
            if (kFma) { fy = __fmaf_rn(px, sn, __fmul_rn(py, cs)); fx = __fmaf_rn(px, cs, -__fmul_rn(py, sn)); }
            else { fy = __fadd_rn(__fmul_rn(px, sn), __fmul_rn(py, cs)); fx = __fsub_rn(__fmul_rn(px, cs), __fmul_rn(py, sn)); }
            tv[e] = cb[__float2int_rn(fy) * pitch + __float2int_rn(fx)];
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

// ------------------------------------------------------------------------------------------------------
// K4 + K6 + K7b, batched (round 2): one warp per GROUP of kOdG keypoint slots of one (image, level) instead of one warp per keypoint.
// What a keypoint-per-warp kernel pays once per keypoint, the group pays once per lane or once per warp:
//   * atan2 + the glibc-exact sincos (about 105 warp-instructions, 18 % of the kernel) run with lane l = keypoint l after the
//     moments of the whole group are known, i.e. once per 16 keypoints instead of once per keypoint with 32 identical lanes;
//   * the lane's 32 pattern coordinates are converted int8 -> float ONCE and stay in registers (no I2F in the sample loop), and
//     cvRound is the add of 1.5 * 2^23 (round-to-nearest-even, exact for |v| < 2^22) whose bias is folded into the patch base
//     address (no F2I either): the sample loop has no instruction left on the quarter-rate conversion pipe;
//   * the disc masks of IC_Angle depend on the lane (= patch row) only and are built once.
// Patches: a ring of kOdRing shared-memory slots per warp, filled by TMA tensor copies on the warp's own mbarriers; request k of a
// group of n keypoints is the orientation patch of keypoint k (k < n) or the descriptor patch of keypoint k - n, so the first
// descriptor patches are in flight while the last orientations are computed and while the lanes evaluate atan2 / sincos.
// Results are bit-identical to k_orient_desc (same integer moments, same separately rounded float operations).
// MEASURED (B200, 320 images, 482 k keypoints): 219 M warp-instructions instead of 285 M, but 0.340 ms (ring of 4 slots, 16 warps per SM;
// 0.43 ms with 3 slots / 20 warps, 0.38 ms with 5 / 12) against 0.327 ms for k_orient_desc: neither kernel is bound by instruction
// issue.  Per keypoint 3.9 KB of patches cross the L2 -> SM path (2.4 GB per launch = 7.1 TB/s, L2 65 % busy, the L1 / shared-memory
// pipe 85 %: TMA fills + the byte gathers of the descriptor, 4.7 wavefronts per load), and a patch ring deep enough to hide the TMA
// latency leaves this kernel shared memory for 16 warps per SM only.  It is therefore OPT-IN (ORBX_DESC_BATCH=1, ORBX_DESC_RING=3..5;
// tests/test_gpu_extract.py::test_batched_orientation_descriptor_kernel checks its bits); the default stays one warp per keypoint.
// ------------------------------------------------------------------------------------------------------
constexpr int kOdG = 16, kOdWarps = 4;
constexpr int kOdSlotBytes = (kDescBytes + 127) & ~127;          // a slot holds either patch (orientation 48 x 31, descriptor 64 x 37)
constexpr int kOdSlotsPerCta = kOdG * kOdWarps;                  // block_info2[blockIdx.x] = level | first slot << 4

constexpr int kOdRedBytes = kOdG * 33 * 4;                       // per warp: partial moments [keypoint][33] (one word per patch row)
__host__ __device__ constexpr int od_smem_bytes(int ring) { return kOdWarps * (ring * kOdSlotBytes + kOdRedBytes); }

template <bool kFma, int kOdRing>
__global__ void __launch_bounds__(kOdWarps * 32, kOdRing <= 3 ? 6 : 5) k_orient_desc_batch(Geom G, const signed char *__restrict__ pattern,
                                                                        const uint32_t *__restrict__ lvl_kp, const int *__restrict__ lvl_n,
                                                                        const int *__restrict__ slot, float *__restrict__ lvl_angle,
                                                                        orbx_kp *__restrict__ out_kp, uint8_t *__restrict__ out_desc,
                                                                        const PatchMaps *__restrict__ maps, const int *__restrict__ block_info,
                                                                        int img0) {
    extern __shared__ __align__(128) uint8_t s_ring[];           // [warp][ring slot][kOdSlotBytes], then [warp][kOdG][33] u32
    __shared__ __align__(8) uint64_t s_bar[kOdWarps * kOdRing];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y + img0;
    const int info = __ldg(block_info + blockIdx.x);
    const int lvl = info & 15;
    const LevelGeom &L = G.lv[lvl];
    const int i0 = (info >> 4) + wid * kOdG;
    const int n = min(kOdG, lvl_n[b * G.nlevels + lvl] - i0);   // keypoints of this warp's group
    if (n <= 0) return;
    const long long first = (long long)b * G.kp_per_image + L.kp_off + i0;
    // lane l < n owns keypoint l of the group: coordinates, output slot, later its moments / angle / rotation
    const uint32_t kp_m = lane < n ? lvl_kp[first + lane] : 0u;
    const int slot_m = lane < n ? slot[first + lane] : 0;
    const int x_m = cand_x(kp_m), y_m = cand_y(kp_m);
    const uint32_t ring = smem_u32(s_ring) + wid * (kOdRing * kOdSlotBytes);
    const uint32_t red = smem_u32(s_ring) + kOdWarps * (kOdRing * kOdSlotBytes) + wid * kOdRedBytes;
    uint64_t *bar = s_bar + wid * kOdRing;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kOdRing; r++) mbar_init(&bar[r], 1);
    }
    __syncwarp();
    const CUtensorMap *map_o = &maps->pyr[lvl], *map_d = &maps->blur[lvl];
    const int nreq = 2 * n;
    int k_issue = 0, s_issue = 0;                                // next request and its ring slot
    // (all lanes call: the coordinates come from the owning lane; lane 0 issues the copy)
    auto issue = [&]() {
        const int j = k_issue < n ? k_issue : k_issue - n;
        const int xj = __shfl_sync(0xFFFFFFFFu, x_m, j), yj = __shfl_sync(0xFFFFFFFFu, y_m, j);
        if (lane == 0) {
            uint64_t *br = &bar[s_issue];
            const uint32_t dst = ring + s_issue * kOdSlotBytes;
            const bool ori = k_issue < n;
            mbar_arrive_expect_tx(br, ori ? (uint32_t)kOriBytes : (uint32_t)kDescBytes);
            const int c0 = ori ? ((kXO + xj - 16) & ~15) : ((kXO + xj - 24) & ~15), c1 = ori ? kYO + yj - kHalfPatch : kYO + yj - 18;
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                         "l"(reinterpret_cast<uint64_t>(ori ? map_o : map_d)), "r"(smem_u32(br)), "r"(c0), "r"(c1), "r"(b)
                         : "memory");
        }
        k_issue++;
        s_issue = s_issue == kOdRing - 1 ? 0 : s_issue + 1;
    };
#pragma unroll
    for (int r = 0; r < kOdRing; r++)
        if (k_issue < nreq) issue();
    // ---- per-lane constants
    // IC_Angle: lane = row v of the 31 x 31 patch; bytes 15 - m .. 15 + m of the row lie inside the disc (m = umax[|v|])
    const int v = lane - kHalfPatch, av = v < 0 ? -v : v;
    uint32_t msk[8];
    {
        const int m = lane < 31 ? (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15) : -1;   // src/ORBextractor.cc:399-413
        const uint32_t bits = m < 0 ? 0u : ((2u << (2 * m)) - 1u) << (kHalfPatch - m);
#pragma unroll
        for (int j = 0; j < 8; j++) msk[j] = ((((bits >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
    }
    // descriptor: lane = output byte; its 8 tests x 2 points as floats (x, y)
    float fpx[16], fpy[16];
    {
        const uint4 pa = __ldg(reinterpret_cast<const uint4 *>(pattern) + 2 * lane), pb = __ldg(reinterpret_cast<const uint4 *>(pattern) + 2 * lane + 1);
        const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};   // word k = (x0, y0, x1, y1) of test k
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                fpx[2 * k + e] = (float)(signed char)(pw[k] >> (16 * e));
                fpy[2 * k + e] = (float)(signed char)(pw[k] >> (16 * e + 8));
            }
    }
    int s_use = 0;                                               // ring slot of the request being consumed
    uint32_t par = 0;                                            // bit s = parity the next wait on slot s expects
    // ---- phase 1: moments of every keypoint of the group (src/ORBextractor.cc:19-43)
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        const int xj = __shfl_sync(0xFFFFFFFFu, x_m, j);
        mbar_wait(&bar[s_use], (par >> s_use) & 1u);
        par ^= 1u << s_use;
        const uint32_t op = ring + s_use * kOdSlotBytes;
        const int c0 = ((kXO + xj - 16) & 15) + 1;               // byte of column x - 15 in the staged row
        int m10 = 0, rowsum = 0;
        ic_row(op + min(lane, 30) * kOriW, c0, msk, m10, rowsum);
        // the row's two sums as one word (|m10| < 2^17, rowsum < 2^13); they are added up after the loop, a keypoint per lane:
        // no shuffle chain inside the loop
        sts32(red + 4 * (j * 33 + lane), ((uint32_t)m10 << 13) | (uint32_t)rowsum);
        __syncwarp();                                            // every lane has read the slot
        if (k_issue < nreq) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue();
        }
        s_use = s_use == kOdRing - 1 ? 0 : s_use + 1;
    }
    // ---- lane l: angle and rotation of keypoint l, its KeyPoint record
    float sn_m = 0.f, cs_m = 1.f;
    __syncwarp();
    if (lane < n) {
        int mm10 = 0, mm01 = 0;
#pragma unroll
        for (int r = 0; r < 31; r++) {                           // (column reads of a [.][33] array: conflict-free)
            const uint32_t pk = lds32(red + 4 * (lane * 33 + r));
            mm10 += (int)pk >> 13;
            mm01 += (r - kHalfPatch) * (int)(pk & 0x1FFFu);
        }
        const float angle = orbx_fast_atan2((float)mm01, (float)mm10);
        const float rad = __fmul_rn(angle, 0.0174532924f);       // (float)(CV_PI/180.f)
        orbx_sincosf(rad, &sn_m, &cs_m);
        lvl_angle[first + lane] = angle;
        orbx_kp o;
        o.x = (float)x_m; o.y = (float)y_m;
        if (lvl != 0) { o.x = __fmul_rn(o.x, L.scale); o.y = __fmul_rn(o.y, L.scale); }
        o.size = L.patch; o.angle = angle; o.response = (float)cand_score(kp_m); o.octave = lvl; o.class_id = -1;
        out_kp[(long long)b * G.out_cap + slot_m] = o;
    }
    // ---- phase 2: 256-bit rBRIEF of every keypoint (src/ORBextractor.cc:46-90)
    constexpr float kRound = 12582912.f;                         // 1.5 * 2^23: (v + kRound) has cvRound(v) in its low mantissa bits
    constexpr uint32_t kRoundBits = 0x4B400000u;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        const int xj = __shfl_sync(0xFFFFFFFFu, x_m, j), sj = __shfl_sync(0xFFFFFFFFu, slot_m, j);
        const float sn = __shfl_sync(0xFFFFFFFFu, sn_m, j), cs = __shfl_sync(0xFFFFFFFFu, cs_m, j);
        mbar_wait(&bar[s_use], (par >> s_use) & 1u);
        par ^= 1u << s_use;
        // shared address of the patch centre, minus the bias of the two rounded coordinates: address = cb + iy * kDescW + ix
        const uint32_t cb = ring + s_use * kOdSlotBytes + 18 * kDescW + 24 + ((kXO + xj - 24) & 15) - kRoundBits * (uint32_t)(kDescW + 1);
        uint32_t val = 0;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            int tv[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float px = fpx[2 * k + e], py = fpy[2 * k + e];
                float fy, fx;
                if (kFma) { fy = __fmaf_rn(px, sn, __fmul_rn(py, cs)); fx = __fmaf_rn(px, cs, -__fmul_rn(py, sn)); }
                else { fy = __fadd_rn(__fmul_rn(px, sn), __fmul_rn(py, cs)); fx = __fsub_rn(__fmul_rn(px, cs), __fmul_rn(py, sn)); }
                const uint32_t iy = __float_as_uint(__fadd_rn(fy, kRound)), ix = __float_as_uint(__fadd_rn(fx, kRound));
                tv[e] = (int)lds8(cb + iy * (uint32_t)kDescW + ix);
            }
            val = __funnelshift_l((uint32_t)(tv[0] - tv[1]), val, 1);   // (val << 1) | (tv[0] < tv[1]): test k ends up in bit k
        }
        out_desc[((long long)b * G.out_cap + sj) * 32 + lane] = (uint8_t)val;
        __syncwarp();
        if (k_issue < nreq) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue();
        }
        s_use = s_use == kOdRing - 1 ? 0 : s_use + 1;
    }
}

}  // namespace orbx
