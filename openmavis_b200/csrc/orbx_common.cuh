// orbx_common.cuh — shared declarations of liborbx (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/orbx.h"

namespace orbx {

constexpr int kEdge = 19;          // EDGE_THRESHOLD, src/ORBextractor.cc:17
constexpr int kXO = 32;            // column of the level interior inside a plane row (>= kEdge, 16-B aligned)
constexpr int kYO = kEdge;         // row of the level interior inside a plane
constexpr int kMaxLevels = 12;
constexpr int kHalfPatch = 15;     // HALF_PATCH_SIZE
constexpr int kFastBorder = 16;    // EDGE_THRESHOLD - 3: minBorderX/Y of ComputeKeyPointsOctTree

// Geometry of one pyramid level, identical for every image of a batch handle.
struct LevelGeom {
    int w, h;                  // interior size  (cvRound(W * inv_scale), src/ORBextractor.cc:1076)
    int pitch, rows;           // plane row pitch in bytes (multiple of 128) and rows (h + 38)
    long long plane_off;       // byte offset of image 0's plane inside the pyramid / blur arena
    long long plane_bytes;     // bytes per image plane
    int tap_x, tap_y;          // offsets of this level's resize taps inside the tap table (int4 entries)
    // FAST cell grid (src/ORBextractor.cc:709-741)
    int n_cols, n_rows, w_cell, h_cell, max_bx, max_by;
    int cell_base;             // first cell of this level in the per-image cell numbering
    int fast_tile_base;        // first FAST tile (CTA) of this level in the per-image tile numbering
    int fast_tiles_x;          // tiles per cell row = ceil(n_cols / Geom::fast_g)
    int cell_cap;              // candidate slots per cell: ceil(w_cell/2)*ceil(h_cell/2) (3x3 strict NMS bound)
    long long cand_off;        // offset (u32 slots) of this level's cell slots / linear list inside an image's arena
    int cand_cap;              // n_cells * cell_cap
    // quadtree roots (src/ORBextractor.cc:505-524)
    int n_ini;
    float h_x;
    // per-level keypoint arena
    int kp_off, kp_cap;
    float scale;               // mvScaleFactor[level]
    float patch;               // (float)(int)(31 * mvScaleFactor[level])
};

struct Geom {
    int nlevels;
    int width, height;
    int total_cells;           // cells per image over all levels
    long long cand_per_image;  // u32 slots per image
    int kp_per_image;          // sum of kp_cap
    int out_cap;               // rows of the final per-image outputs
    int fma_mode;
    int fast_g;                // cells per FAST tile
    int fast_ip, fast_ymax;    // shared-memory row pitch (bytes, multiple of 32) and max tested rows of a FAST tile
    int fast_total_tiles;      // FAST tiles per image over all levels
    LevelGeom lv[kMaxLevels];
};

#ifdef __CUDACC__
__device__ __forceinline__ int hamming256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    // ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2577-2591) == popcount(a ^ b) over 8 words
    return (__popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z)) +
           (__popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y)) + (__popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w));
}

// Same distance with 6 POPC instead of 8: two carry-save adders (3 words -> ones + twos, 2 LOP3 each) move a quarter of the
// popcount work from the XU pipe (POPC: 16 lanes / clk / SM, the bound of the brute-force matcher) to the 4x wider ALU pipe.
__device__ __forceinline__ int hamming256_csa(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    const uint32_t sA = x0 ^ x1 ^ x2, cA = (x0 & x1) | (x2 & (x0 | x1));
    const uint32_t sB = x3 ^ x4 ^ x5, cB = (x3 & x4) | (x5 & (x3 | x4));
    return (__popc(sA) + __popc(sB)) + (__popc(x6) + __popc(x7)) + 2 * (__popc(cA) + __popc(cB));
}
#endif

// Thread-local error string behind orbx_last_error().
void set_error(const std::string &msg);
int fail(orbx_status st, const std::string &msg);

#define ORBX_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::orbx::fail(ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// mnFeaturesPerLevel / scale tables computed exactly like the constructor (src/ORBextractor.cc:351-391).
void level_scale_tables(float scale_factor, int nlevels, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2);
void features_per_level(int nfeatures, float scale_factor, int nlevels, int *out);

}  // namespace orbx
