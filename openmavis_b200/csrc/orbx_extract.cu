// orbx_extract.cu — host side of the extraction C ABI (include/orbx.h): geometry tables, device arenas,
// launch sequence.  Replaces ORBextractor::ORBextractor / operator() (src/ORBextractor.cc:351-414, 987-1071).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/orbx_pattern.h"
#include "orbx_extract_kernels.cuh"
#include "orbx_stereo_kernels.cuh"

namespace orbx {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(orbx_status st, const std::string &msg) { g_err = msg; return (int)st; }
const char *last_error() { return g_err.c_str(); }

// src/ORBextractor.cc:362-376
void level_scale_tables(float scale_factor, int nlevels, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2) {
    const double scale = scale_factor;   // `double scaleFactor` member (include/ORBextractor.h:70)
    std::vector<float> s(nlevels), g(nlevels);
    s[0] = 1.0f; g[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { s[i] = (float)(s[i - 1] * scale); g[i] = s[i] * s[i]; }
    for (int i = 0; i < nlevels; i++) {
        if (sf) sf[i] = s[i];
        if (sigma2) sigma2[i] = g[i];
        if (inv_sf) inv_sf[i] = 1.0f / s[i];
        if (inv_sigma2) inv_sigma2[i] = 1.0f / g[i];
    }
}

// src/ORBextractor.cc:378-391
void features_per_level(int nfeatures, float scale_factor, int nlevels, int *out) {
    const double scale = scale_factor;
    float factor = (float)(1.0f / scale);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        out[l] = (int)lrintf(desired);
        sum += out[l];
        desired *= factor;
    }
    out[nlevels - 1] = std::max(nfeatures - sum, 0);
}

}  // namespace orbx

using namespace orbx;

struct orbx_extractor {
    int device = 0;
    int width = 0, height = 0, max_images = 0, max_nfeatures = 0, nlevels = 0;
    float scale_factor = 1.2f;
    Geom G;
    BlurTiles blur_tiles;
    BlurMaps blur_maps;
    FastMaps fast_maps;
    PatchMaps patch_maps;
    PatchMaps *d_patch_maps = nullptr;
    ResizeMap rsz_maps[kMaxLevels];                  // [l] = source level l - 1 of the resize that produces level l
    int blur_total_tiles = 0;
    size_t fast_smem = 0, oct_smem = 0, rsz_smem = 0;
    int rsz_rows = 0, rsz_pitch = 0;
    int oct_list_cap = 0, oct_threads = 256;
    // small chunks (one rig frame) are launch-bound: their ~15 launches and the two-stream fork / join are captured once per
    // (lane, first image, count, pixels) into a CUDA graph and replayed with one launch
    struct ChunkGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };
    std::map<std::tuple<int, int, int, const void *>, ChunkGraph> graphs;
    bool use_graphs = true;
    int oct_wide_max = 148;              // (image, level) pairs up to which the quadtree runs as 1024-thread CTAs (one per SM)
    long long pyr_bytes_per_batch = 0;
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    // chunk pipelines: a batch is cut into chunks of chunk_images; chunk k runs on lane k % kLanes (main + aux stream: the
    // blur of a chunk overlaps its FAST / quadtree), so the latency-bound stages of one chunk hide behind the
    // bandwidth-bound stages of the next; host-buffer calls add a copy-in and a copy-out stream.
    static constexpr int kLanes = 2;
    cudaStream_t lane_main[kLanes] = {}, lane_aux[kLanes] = {}, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_pyr[kLanes] = {}, ev_blur[kLanes] = {}, ev_lane[kLanes] = {}, ev_d2h = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out;
    int chunk_images = 0;                            // 0 = choose from the batch size
    static constexpr int kRing = 16;                 // profiled calls kept for orbx_extractor_stage_ms
    cudaEvent_t ev[kRing][ORBX_NUM_STAGES] = {};
    int prof_calls = 0;                              // profiled calls since profiling was switched on
    cudaEvent_t ev_staging = nullptr;
    bool profiling = false, have_times = false;
    int last_launches = 0, last_n = 0;

    // device
    uint8_t *d_in = nullptr, *d_pyr = nullptr, *d_blur = nullptr, *d_out_desc = nullptr;
    ResizeTap *d_taps = nullptr;
    signed char *d_pattern = nullptr;
    uint32_t *d_cells = nullptr, *d_lin = nullptr, *d_lvl_kp = nullptr;
    unsigned short *d_node = nullptr;
    int *d_cell_cnt = nullptr, *d_cell_off = nullptr, *d_lin_n = nullptr, *d_lvl_n = nullptr, *d_slot = nullptr;
    int *d_n_kp = nullptr, *d_n_mono = nullptr, *d_nfeat = nullptr, *d_desc_blocks = nullptr, *d_blur_tiles = nullptr;
    int desc_blocks = 0;
    int *d_desc_blocks2 = nullptr;   // k_orient_desc_batch: CTA -> level | first slot << 4 (kOdSlotsPerCta slots per CTA)
    int desc_blocks2 = 0;
    bool desc_batch = false;         // k_orient_desc_batch instead of k_orient_desc (measured equal to slower: see the kernel's comment)
    int desc_ring = 4;
    float *d_lvl_angle = nullptr;
    orbx_kp *d_out_kp = nullptr;
    orbx_image_params *d_params = nullptr;
    float *d_st_uright = nullptr, *d_st_depth = nullptr;   // ComputeStereoMatches outputs (allocated on first use)
    int *d_st_sad = nullptr, *d_st_kept = nullptr;
    FastTile *d_fast_tiles = nullptr;
    int fast_kb_words = 0;
    FastStats *d_fast_stats = nullptr;                     // k_fast counters, accumulated while profiling is on
    // pinned staging
    uint8_t *h_in = nullptr;
    orbx_image_params *h_params = nullptr;
    int *h_nfeat = nullptr;
};

// k_fast is compiled for the shared-memory row pitches a geometry can ask for: 16 x odd bytes up to 336
typedef void (*fast_kernel_t)(const uint8_t *, const FastTile *, const FastArgs, const orbx_image_params *, uint32_t *, int *, const FastMaps, FastStats *);
static const int kFastPitches[] = {80, 112, 144, 176, 208, 240, 272, 304, 336};
static int fast_pitch_for(int need) {
    for (int p : kFastPitches) if (p >= need) return p;
    return 0;
}
static fast_kernel_t fast_kernel(int ip) {
    switch (ip) {
        case 80: return k_fast<80>;
        case 112: return k_fast<112>;
        case 144: return k_fast<144>;
        case 176: return k_fast<176>;
        case 208: return k_fast<208>;
        case 240: return k_fast<240>;
        case 272: return k_fast<272>;
        case 304: return k_fast<304>;
        default: return k_fast<336>;
    }
}

static int build_geometry(orbx_extractor *ex, int fma_mode) {
    Geom &G = ex->G;
    memset(&G, 0, sizeof(G));
    const int nl = ex->nlevels;
    G.nlevels = nl; G.width = ex->width; G.height = ex->height; G.fma_mode = fma_mode;
    ex->sf.resize(nl); ex->inv_sf.resize(nl); ex->sigma2.resize(nl); ex->inv_sigma2.resize(nl);
    level_scale_tables(ex->scale_factor, nl, ex->sf.data(), ex->inv_sf.data(), ex->sigma2.data(), ex->inv_sigma2.data());
    std::vector<int> nfeat(nl);
    features_per_level(ex->max_nfeatures, ex->scale_factor, nl, nfeat.data());
    long long plane_off = 0, cand_off = 0;
    int cell_base = 0, kp_off = 0, tap_off = 0, max_cell_w = 0, max_cell_h = 0, max_list = 0;
    for (int l = 0; l < nl; l++) {
        LevelGeom &L = G.lv[l];
        L.w = (int)lrintf((float)ex->width * ex->inv_sf[l]);      // src/ORBextractor.cc:1075-1076
        L.h = (int)lrintf((float)ex->height * ex->inv_sf[l]);
        if (L.w < 2 * kFastBorder + 36 || L.h < 2 * kFastBorder + 36)
            return fail(ORBX_ERR_INVALID, "image too small for this pyramid depth (level " + std::to_string(l) + ")");
        L.pitch = (kXO + L.w + kEdge + 127) & ~127;
        L.rows = L.h + 2 * kEdge;
        L.plane_bytes = (long long)L.pitch * L.rows;
        L.plane_off = plane_off;
        plane_off += L.plane_bytes * ex->max_images;
        L.tap_x = tap_off; tap_off += L.w;
        L.tap_y = tap_off; tap_off += L.h;
        // src/ORBextractor.cc:709-724
        L.max_bx = L.w - kEdge + 3; L.max_by = L.h - kEdge + 3;
        const float width = (float)(L.max_bx - kFastBorder), height = (float)(L.max_by - kFastBorder);
        L.n_cols = (int)(width / 35.f); L.n_rows = (int)(height / 35.f);
        L.w_cell = (int)ceilf(width / L.n_cols); L.h_cell = (int)ceilf(height / L.n_rows);
        L.cell_base = cell_base; cell_base += L.n_cols * L.n_rows;
        L.cell_cap = ((L.w_cell + 1) / 2) * ((L.h_cell + 1) / 2);
        L.cand_off = cand_off;
        L.cand_cap = L.n_cols * L.n_rows * L.cell_cap;
        cand_off += L.cand_cap;
        max_cell_w = std::max(max_cell_w, L.w_cell + 6); max_cell_h = std::max(max_cell_h, L.h_cell + 6);
        // src/ORBextractor.cc:505-507
        L.n_ini = (int)roundf((float)(L.max_bx - kFastBorder) / (L.max_by - kFastBorder));
        if (L.n_ini < 1) return fail(ORBX_ERR_INVALID, "image aspect ratio below 1:2 is not supported (nIni == 0)");
        L.h_x = (float)(L.max_bx - kFastBorder) / L.n_ini;
        const int lcap = oct_list_cap(nfeat[l], L.n_ini) + 8;
        max_list = std::max(max_list, lcap);
        L.kp_off = kp_off; L.kp_cap = lcap; kp_off += lcap;
        L.scale = ex->sf[l];
        L.patch = (float)(int)(31 * ex->sf[l]);                   // src/ORBextractor.cc:806
        if (L.w + 2 * kFastBorder >= 4096 || L.h + 2 * kFastBorder >= 4096)
            return fail(ORBX_ERR_INVALID, "images larger than 4000 px are not supported");
    }
    G.total_cells = cell_base; G.cand_per_image = cand_off; G.kp_per_image = kp_off; G.out_cap = kp_off;
    ex->pyr_bytes_per_batch = plane_off;
    // FAST tiles: up to kFastMaxG adjacent cells of a cell row per CTA (one warp per cell).  The shared-memory row pitch is a template
    // constant of k_fast: 16 x odd bytes (strips of 7 rows then start in different banks), sized for level 0's cells; levels whose
    // cells are wider (ceil() makes the small levels' cells a few pixels wider) put fewer cells into a tile.
    int want_g = kFastMaxG;
    if (const char *t = getenv("ORBX_FAST_G")) want_g = std::max(1, std::min(want_g, atoi(t)));   // tuning knob
    G.fast_g = want_g;
    G.fast_ip = fast_pitch_for(want_g * G.lv[0].w_cell + 36);
    if (G.fast_ip == 0) { G.fast_g = 1; G.fast_ip = fast_pitch_for(max_cell_w - 6 + 36); }
    if (G.fast_ip == 0) return fail(ORBX_ERR_INVALID, "FAST cells wider than 300 px are not supported");
    int ftiles = 0;
    for (int l = 0; l < nl; l++) {
        LevelGeom &L = G.lv[l];
        const int g = std::max(1, std::min(G.fast_g, (G.fast_ip - 36) / L.w_cell));
        if (L.w_cell + 36 > G.fast_ip) return fail(ORBX_ERR_INVALID, "FAST cell does not fit the tile pitch");
        L.fast_tiles_x = (L.n_cols + g - 1) / g;
        L.fast_tile_base = ftiles;
        ftiles += L.fast_tiles_x * L.n_rows;
    }
    G.fast_total_tiles = ftiles;
    G.fast_ymax = max_cell_h - 6;
    {   // keep bitmap of one cell: one bit per tested pixel, 64 bits per lane
        int bits = 0;
        for (int l = 0; l < nl; l++) bits = std::max(bits, G.lv[l].w_cell * G.lv[l].h_cell);
        ex->fast_kb_words = ((bits + 63) / 64 * 2 + 3) & ~3;
    }
    ex->fast_smem = fast_smem_bytes(G.fast_ip, G.fast_ymax, ex->fast_kb_words);
    if (ex->fast_smem > 200 * 1024 || G.fast_ip > 512 || (G.fast_ymax + 6) * G.fast_ip > 65535)
        return fail(ORBX_ERR_INVALID, "FAST tile does not fit in shared memory");
    ex->oct_list_cap = max_list;
    ex->oct_smem = oct_work_bytes(max_list);
    if (ex->oct_smem > 200 * 1024) return fail(ORBX_ERR_INVALID, "nfeatures too large for the quadtree workspace");
    if (const char *t = getenv("ORBX_OCT_SMEM_KB")) ex->oct_smem = std::max(ex->oct_smem, (size_t)atoi(t) * 1024);   // tuning knob: caps CTAs per SM
    int tiles = 0;
    for (int l = 0; l < nl; l++) {
        ex->blur_tiles.base[l] = tiles;
        tiles += ((G.lv[l].w + kBlurTW - 1) / kBlurTW) * ((G.lv[l].h + kBlurTH - 1) / kBlurTH);
    }
    for (int l = nl; l <= kMaxLevels; l++) ex->blur_tiles.base[l] = tiles;
    ex->blur_total_tiles = tiles;
    return ORBX_OK;
}

template <class T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

extern "C" {

const char *orbx_last_error(void) { return orbx::last_error(); }
const char *orbx_version(void) { return "orbx 0.1 (sm_100a)"; }

int orbx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int orbx_features_per_level(int nfeatures, float scale_factor, int nlevels, int *out) {
    if (!out || nlevels < 1 || nlevels > kMaxLevels || nfeatures < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    features_per_level(nfeatures, scale_factor, nlevels, out);
    return ORBX_OK;
}

void orbx_extractor_destroy(orbx_extractor *ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    cudaFree(ex->d_in); cudaFree(ex->d_pyr); cudaFree(ex->d_blur); cudaFree(ex->d_out_desc); cudaFree(ex->d_taps);
    cudaFree(ex->d_pattern); cudaFree(ex->d_cells); cudaFree(ex->d_lin); cudaFree(ex->d_lvl_kp); cudaFree(ex->d_node);
    cudaFree(ex->d_cell_cnt); cudaFree(ex->d_cell_off); cudaFree(ex->d_lin_n); cudaFree(ex->d_lvl_n); cudaFree(ex->d_slot);
    cudaFree(ex->d_n_kp); cudaFree(ex->d_n_mono); cudaFree(ex->d_nfeat); cudaFree(ex->d_lvl_angle); cudaFree(ex->d_out_kp);
    cudaFree(ex->d_params); cudaFree(ex->d_fast_stats); cudaFree(ex->d_fast_tiles); cudaFree(ex->d_patch_maps); cudaFree(ex->d_desc_blocks); cudaFree(ex->d_desc_blocks2); cudaFree(ex->d_blur_tiles);
    cudaFree(ex->d_st_uright); cudaFree(ex->d_st_depth); cudaFree(ex->d_st_sad); cudaFree(ex->d_st_kept);
    if (ex->h_in) cudaFreeHost(ex->h_in);
    if (ex->h_params) cudaFreeHost(ex->h_params);
    if (ex->h_nfeat) cudaFreeHost(ex->h_nfeat);
    for (auto &g : ex->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    for (auto &r : ex->ev) for (auto &e : r) if (e) cudaEventDestroy(e);
    if (ex->ev_staging) cudaEventDestroy(ex->ev_staging);
    for (int i = 0; i < orbx_extractor::kLanes; i++) {
        if (ex->lane_main[i]) cudaStreamDestroy(ex->lane_main[i]);
        if (ex->lane_aux[i]) cudaStreamDestroy(ex->lane_aux[i]);
        if (ex->ev_pyr[i]) cudaEventDestroy(ex->ev_pyr[i]);
        if (ex->ev_blur[i]) cudaEventDestroy(ex->ev_blur[i]);
        if (ex->ev_lane[i]) cudaEventDestroy(ex->ev_lane[i]);
    }
    if (ex->s_h2d) cudaStreamDestroy(ex->s_h2d);
    if (ex->s_d2h) cudaStreamDestroy(ex->s_d2h);
    if (ex->ev_fork) cudaEventDestroy(ex->ev_fork);
    if (ex->ev_d2h) cudaEventDestroy(ex->ev_d2h);
    for (auto &v : ex->ev_in) cudaEventDestroy(v);
    for (auto &v : ex->ev_out) cudaEventDestroy(v);
    if (ex->own_stream) cudaStreamDestroy(ex->own_stream);
    delete ex;
}

int orbx_extractor_create(orbx_extractor **out, int device, int width, int height, int max_images, int max_nfeatures,
                          float scale_factor, int nlevels, int fma_mode) {
    if (!out) return fail(ORBX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (width <= 0 || height <= 0 || max_images <= 0 || max_nfeatures <= 0 || nlevels < 1 || nlevels > kMaxLevels ||
        !(scale_factor > 1.0f))
        return fail(ORBX_ERR_INVALID, "bad extractor arguments");
    int ndev = orbx_device_count();
    if (ndev <= 0) return fail(ORBX_ERR_NO_DEVICE, "no CUDA device (liborbx has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ORBX_ERR_NO_DEVICE, "device index out of range");
    orbx_extractor *ex = new orbx_extractor();
    ex->device = device; ex->width = width; ex->height = height; ex->max_images = max_images;
    ex->max_nfeatures = max_nfeatures; ex->nlevels = nlevels; ex->scale_factor = scale_factor;
    if (const char *t = getenv("ORBX_OCT_WIDE")) ex->oct_wide_max = atoi(t);   // tuning knob
    if (const char *t = getenv("ORBX_NO_GRAPH")) ex->use_graphs = atoi(t) == 0;
    if (const char *t = getenv("ORBX_OCT_THREADS")) ex->oct_threads = std::max(32, std::min(256, atoi(t) & ~31));   // tuning knob
    int st = build_geometry(ex, fma_mode ? 1 : 0);
    if (st != ORBX_OK) { delete ex; return st; }
    const Geom &G = ex->G;
    const size_t B = (size_t)max_images;
#define EX_CUDA(call)                                                                                 \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            std::string m = std::string(#call) + ": " + cudaGetErrorString(_e);                       \
            orbx_extractor_destroy(ex);                                                               \
            return fail(ORBX_ERR_CUDA, m);                                                            \
        }                                                                                             \
    } while (0)
    EX_CUDA(cudaSetDevice(device));
    EX_CUDA(cudaStreamCreateWithFlags(&ex->own_stream, cudaStreamNonBlocking));
    ex->stream = ex->own_stream;
    for (auto &r : ex->ev) for (auto &e : r) EX_CUDA(cudaEventCreate(&e));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_staging, cudaEventDisableTiming));
    for (int i = 0; i < orbx_extractor::kLanes; i++) {
        EX_CUDA(cudaStreamCreateWithFlags(&ex->lane_main[i], cudaStreamNonBlocking));
        EX_CUDA(cudaStreamCreateWithFlags(&ex->lane_aux[i], cudaStreamNonBlocking));
        EX_CUDA(cudaEventCreateWithFlags(&ex->ev_pyr[i], cudaEventDisableTiming));
        EX_CUDA(cudaEventCreateWithFlags(&ex->ev_blur[i], cudaEventDisableTiming));
        EX_CUDA(cudaEventCreateWithFlags(&ex->ev_lane[i], cudaEventDisableTiming));
    }
    EX_CUDA(cudaStreamCreateWithFlags(&ex->s_h2d, cudaStreamNonBlocking));
    EX_CUDA(cudaStreamCreateWithFlags(&ex->s_d2h, cudaStreamNonBlocking));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_fork, cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&ex->ev_d2h, cudaEventDisableTiming));
    if (const char *t = getenv("ORBX_CHUNK")) ex->chunk_images = std::max(0, atoi(t));   // tuning knob
    EX_CUDA(dalloc(&ex->d_in, B * width * height));
    EX_CUDA(dalloc(&ex->d_pyr, (size_t)ex->pyr_bytes_per_batch));
    EX_CUDA(dalloc(&ex->d_blur, (size_t)ex->pyr_bytes_per_batch));
    EX_CUDA(dalloc(&ex->d_cells, B * G.cand_per_image));
    EX_CUDA(dalloc(&ex->d_lin, B * G.cand_per_image));
    EX_CUDA(dalloc(&ex->d_node, B * G.cand_per_image));
    EX_CUDA(dalloc(&ex->d_cell_cnt, B * G.total_cells));
    EX_CUDA(dalloc(&ex->d_cell_off, B * G.total_cells));
    EX_CUDA(dalloc(&ex->d_lin_n, B * nlevels));
    EX_CUDA(dalloc(&ex->d_lvl_n, B * nlevels));
    EX_CUDA(dalloc(&ex->d_nfeat, B * nlevels));
    EX_CUDA(dalloc(&ex->d_lvl_kp, B * G.kp_per_image));
    EX_CUDA(dalloc(&ex->d_lvl_angle, B * G.kp_per_image));
    EX_CUDA(dalloc(&ex->d_slot, B * G.kp_per_image));
    EX_CUDA(dalloc(&ex->d_out_kp, B * G.out_cap));
    EX_CUDA(dalloc(&ex->d_out_desc, B * G.out_cap * 32));
    EX_CUDA(dalloc(&ex->d_n_kp, B));
    EX_CUDA(dalloc(&ex->d_n_mono, B));
    EX_CUDA(dalloc(&ex->d_params, B));
    EX_CUDA(cudaMallocHost((void **)&ex->h_in, B * width * height));
    EX_CUDA(cudaMallocHost((void **)&ex->h_params, B * sizeof(orbx_image_params)));
    EX_CUDA(cudaMallocHost((void **)&ex->h_nfeat, B * nlevels * sizeof(int)));
    EX_CUDA(cudaMemsetAsync(ex->d_pyr, 0, (size_t)ex->pyr_bytes_per_batch, ex->stream));
    EX_CUDA(cudaMemsetAsync(ex->d_blur, 0, (size_t)ex->pyr_bytes_per_batch, ex->stream));
    EX_CUDA(cudaMemsetAsync(ex->d_n_kp, 0, B * sizeof(int), ex->stream));
    EX_CUDA(cudaMemsetAsync(ex->d_n_mono, 0, B * sizeof(int), ex->stream));
    EX_CUDA(cudaMemsetAsync(ex->d_lvl_n, 0, B * nlevels * sizeof(int), ex->stream));
    EX_CUDA(cudaMemsetAsync(ex->d_lin_n, 0, B * nlevels * sizeof(int), ex->stream));
    // resize taps
    {
        int ntap = 0;
        for (int l = 0; l < nlevels; l++) ntap += G.lv[l].w + G.lv[l].h;
        std::vector<ResizeTap> taps(ntap);
        for (int l = 1; l < nlevels; l++) {
            orbx_resize_taps(G.lv[l - 1].w, G.lv[l].w, taps.data() + G.lv[l].tap_x);
            orbx_resize_taps(G.lv[l - 1].h, G.lv[l].h, taps.data() + G.lv[l].tap_y);
        }
        // staged source extent of the largest resize tile (rows, 16-byte aligned columns)
        for (int l = 1; l < nlevels; l++) {
            const ResizeTap *tx = taps.data() + G.lv[l].tap_x, *ty = taps.data() + G.lv[l].tap_y;
            for (int y0 = 0; y0 < G.lv[l].h; y0 += kRszTH)
                ex->rsz_rows = std::max(ex->rsz_rows, ty[std::min(y0 + kRszTH, G.lv[l].h) - 1].s1 - ty[y0].s0 + 1);
            for (int x0 = 0; x0 < G.lv[l].w; x0 += kRszTW)
                ex->rsz_pitch = std::max(ex->rsz_pitch, ((tx[std::min(x0 + kRszTW, G.lv[l].w) - 1].s1 - (tx[x0].s0 & ~15) + 16) >> 4) << 4);
        }
        ex->rsz_smem = resize_smem_bytes(ex->rsz_rows, ex->rsz_pitch);
        if (ex->rsz_smem > 200 * 1024) { orbx_extractor_destroy(ex); return fail(ORBX_ERR_INVALID, "scale factor too large for the resize tile"); }
        EX_CUDA(dalloc(&ex->d_taps, (size_t)ntap));
        EX_CUDA(cudaMemcpy(ex->d_taps, taps.data(), ntap * sizeof(ResizeTap), cudaMemcpyHostToDevice));
        std::vector<FastTile> tinfo;   // FAST tiles (CTAs) of one image: a cell row is cut into equal tiles of at most fast_g cells
        for (int l = 0; l < nlevels; l++) {
            const LevelGeom &L = G.lv[l];
            const int cpt = (L.n_cols + L.fast_tiles_x - 1) / L.fast_tiles_x;
            for (int i = 0; i < L.n_rows; i++)
                for (int j = 0; j < L.fast_tiles_x; j++) {
                    FastTile T;
                    memset(&T, 0, sizeof(T));
                    const int j0 = j * cpt, j1 = std::min(j0 + cpt, L.n_cols);
                    T.lvl = l; T.ncell = j1 - j0; T.wc = L.w_cell; T.cell_cap = L.cell_cap;
                    // tested area of the tile in level coordinates (see the header comment of k_fast)
                    T.tx0 = kEdge + j0 * L.w_cell; T.ty0 = kEdge + i * L.h_cell;
                    T.X = std::max(0, std::min(kEdge + j1 * L.w_cell, L.w - kEdge) - T.tx0);
                    T.Y = std::max(0, std::min(T.ty0 + L.h_cell, L.h - kEdge) - T.ty0);
                    T.pc0 = (kXO + T.tx0 - 3) & ~15; T.prow0 = kYO + T.ty0 - 3; T.c0 = kXO + T.tx0 - T.pc0;
                    T.cell0 = L.cell_base + i * L.n_cols + j0;
                    T.nstr = 1; T.rs = 1;
                    if (T.Y > 0) fast_plan_strips(T.Y, T.wc, G.fast_ip, &T.nstr, &T.rs, T.sr);
                    const long long cand0 = L.cand_off + (long long)(i * L.n_cols + j0) * L.cell_cap;
                    T.cand0_lo = (int)(unsigned)(cand0 & 0xFFFFFFFFll); T.cand0_hi = (int)(cand0 >> 32);
                    T.plane_off = L.plane_off; T.plane_bytes = L.plane_bytes; T.pitch = L.pitch;
                    tinfo.push_back(T);
                }
        }
        if ((int)tinfo.size() != G.fast_total_tiles) { orbx_extractor_destroy(ex); return fail(ORBX_ERR_INVALID, "internal: FAST tile count"); }
        EX_CUDA(dalloc(&ex->d_fast_stats, (size_t)1));
        EX_CUDA(cudaMemset(ex->d_fast_stats, 0, sizeof(FastStats)));
        EX_CUDA(dalloc(&ex->d_fast_tiles, tinfo.size()));
        EX_CUDA(cudaMemcpy(ex->d_fast_tiles, tinfo.data(), tinfo.size() * sizeof(FastTile), cudaMemcpyHostToDevice));
        std::vector<int> btiles;   // blur CTA -> level | tile row << 4 | tile column << 16
        for (int l = 0; l < nlevels; l++)
            for (int ty = 0; ty < (G.lv[l].h + kBlurTH - 1) / kBlurTH; ty++)
                for (int tx = 0; tx < (G.lv[l].w + kBlurTW - 1) / kBlurTW; tx++) btiles.push_back(l | (ty << 4) | (tx << 16));
        if ((int)btiles.size() != ex->blur_total_tiles) { orbx_extractor_destroy(ex); return fail(ORBX_ERR_INVALID, "internal: blur tile count"); }
        EX_CUDA(dalloc(&ex->d_blur_tiles, btiles.size()));
        EX_CUDA(cudaMemcpy(ex->d_blur_tiles, btiles.data(), btiles.size() * sizeof(int), cudaMemcpyHostToDevice));
        std::vector<int> binfo;   // orientation / descriptor CTA -> level | first keypoint slot << 4  (8 slots per CTA)
        for (int l = 0; l < nlevels; l++)
            for (int i0 = 0; i0 < G.lv[l].kp_cap; i0 += 8) binfo.push_back(l | (i0 << 4));
        ex->desc_blocks = (int)binfo.size();
        EX_CUDA(dalloc(&ex->d_desc_blocks, binfo.size()));
        EX_CUDA(cudaMemcpy(ex->d_desc_blocks, binfo.data(), binfo.size() * sizeof(int), cudaMemcpyHostToDevice));
        std::vector<int> binfo2;
        for (int l = 0; l < nlevels; l++)
            for (int i0 = 0; i0 < G.lv[l].kp_cap; i0 += kOdSlotsPerCta) binfo2.push_back(l | (i0 << 4));
        ex->desc_blocks2 = (int)binfo2.size();
        EX_CUDA(dalloc(&ex->d_desc_blocks2, binfo2.size()));
        EX_CUDA(cudaMemcpy(ex->d_desc_blocks2, binfo2.data(), binfo2.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (const char *t = getenv("ORBX_DESC_BATCH")) ex->desc_batch = atoi(t) != 0;   // tuning knob: 1 = one warp per group of 16 keypoints
        if (const char *t = getenv("ORBX_DESC_RING")) ex->desc_ring = std::max(3, std::min(5, atoi(t)));
        static const signed char pat[1024] = {ORBX_PATTERN_VALUES};
        EX_CUDA(dalloc(&ex->d_pattern, (size_t)1024));
        EX_CUDA(cudaMemcpy(ex->d_pattern, pat, 1024, cudaMemcpyHostToDevice));
    }
    {
        // Dynamic shared memory limits.  The attribute is per kernel and per process, not per handle, and the kernels also hold a few
        // bytes of static shared memory: every handle raises the limit to the device's opt-in maximum, so a handle whose tiles stay
        // just under 48 KB, or one created after a handle with larger tiles, never runs into another handle's setting.
        int optin = 0;
        EX_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        if ((size_t)optin < std::max(std::max(ex->fast_smem, ex->oct_smem), ex->rsz_smem) + 1024) {
            orbx_extractor_destroy(ex);
            return fail(ORBX_ERR_INVALID, "the tiles of this geometry do not fit in the device's shared memory");
        }
        EX_CUDA(cudaFuncSetAttribute(k_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        EX_CUDA(cudaFuncSetAttribute(fast_kernel(G.fast_ip), cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        EX_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        EX_CUDA(cudaFuncSetAttribute(k_octree_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        EX_CUDA(cudaFuncSetAttribute(k_blur, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(3)));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(3)));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(4)));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(4)));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(5)));
        EX_CUDA(cudaFuncSetAttribute(k_orient_desc_batch<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, od_smem_bytes(5)));
    }
    {   // TMA tensor maps of the pyramid levels for the blur tiles (driver entry point through the runtime: no -lcuda)
        memset(&ex->blur_maps, 0, sizeof(ex->blur_maps));
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        bool ok = !getenv("ORBX_NO_TMA") && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
                  qres == cudaDriverEntryPointSuccess && fn;
        for (int l = 0; ok && l < nlevels; l++) {
            const LevelGeom &L = G.lv[l];
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)max_images};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.plane_bytes};
            const cuuint32_t box[3] = {(cuuint32_t)kBlurSP, (cuuint32_t)kBlurSR, 1u}, estr[3] = {1u, 1u, 1u};
            ok = ((encode_fn)fn)(&ex->blur_maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_pyr + L.plane_off, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        // FAST tiles are up to 512 bytes wide and a TMA box is at most 256 elements per dimension: the levels are described as
        // tensors of 16-bit elements (x extent and coordinates in pixel pairs)
        bool ok3 = ok && G.fast_ip / 2 <= 256 && G.fast_ymax + 6 <= 256;
        memset(&ex->fast_maps, 0, sizeof(ex->fast_maps));
        for (int l = 0; ok3 && l < nlevels; l++) {
            const LevelGeom &L = G.lv[l];
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch / 2, (cuuint64_t)L.rows, (cuuint64_t)max_images};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.plane_bytes};
            const cuuint32_t box[3] = {(cuuint32_t)G.fast_ip / 2, (cuuint32_t)(G.fast_ymax + 6), 1u}, estr[3] = {1u, 1u, 1u};
            ok3 = ((encode_fn)fn)(&ex->fast_maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, ex->d_pyr + L.plane_off, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        ex->fast_maps.use_tma = ok3 ? 1 : 0;
        bool ok4 = ok;
        memset(&ex->patch_maps, 0, sizeof(ex->patch_maps));
        for (int l = 0; ok4 && l < nlevels; l++) {
            const LevelGeom &L = G.lv[l];
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)max_images};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.plane_bytes};
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            const cuuint32_t box_o[3] = {(cuuint32_t)kOriW, (cuuint32_t)kOriH, 1u}, box_d[3] = {(cuuint32_t)kDescW, (cuuint32_t)kDescH, 1u};
            ok4 = ((encode_fn)fn)(&ex->patch_maps.pyr[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_pyr + L.plane_off, dims, strides, box_o, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                  ((encode_fn)fn)(&ex->patch_maps.blur[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_blur + L.plane_off, dims, strides, box_d, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        ex->patch_maps.use_tma = (ok4 && !getenv("ORBX_NO_PATCH_TMA")) ? 1 : 0;
        EX_CUDA(cudaMalloc((void **)&ex->d_patch_maps, sizeof(PatchMaps)));
        EX_CUDA(cudaMemcpy(ex->d_patch_maps, &ex->patch_maps, sizeof(PatchMaps), cudaMemcpyHostToDevice));
        bool ok2 = ok && ex->rsz_rows <= 256 && ex->rsz_pitch <= 256;
        memset(ex->rsz_maps, 0, sizeof(ex->rsz_maps));
        for (int l = 1; ok2 && l < nlevels; l++) {
            const LevelGeom &L = G.lv[l - 1];
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)max_images};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.plane_bytes};
            const cuuint32_t box[3] = {(cuuint32_t)ex->rsz_pitch, (cuuint32_t)ex->rsz_rows, 1u}, estr[3] = {1u, 1u, 1u};
            ok2 = ((encode_fn)fn)(&ex->rsz_maps[l].m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->d_pyr + L.plane_off, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        for (int l = 1; l < nlevels; l++) ex->rsz_maps[l].use_tma = ok2 ? 1 : 0;
        cudaGetLastError();
        ex->blur_maps.use_tma = ok ? 1 : 0;      // otherwise the kernel stages the tile with cp.async
    }
    EX_CUDA(cudaStreamSynchronize(ex->stream));
#undef EX_CUDA
    *out = ex;
    return ORBX_OK;
}

int orbx_extractor_set_stream(orbx_extractor *ex, void *cuda_stream) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    ex->stream = cuda_stream ? (cudaStream_t)cuda_stream : ex->own_stream;
    return ORBX_OK;
}

int orbx_extractor_capacity(const orbx_extractor *ex) { return ex ? ex->G.out_cap : (int)ORBX_ERR_INVALID; }

int orbx_extractor_tables(const orbx_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                          int *level_w, int *level_h) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    for (int l = 0; l < ex->nlevels; l++) {
        if (scale) scale[l] = ex->sf[l];
        if (inv_scale) inv_scale[l] = ex->inv_sf[l];
        if (sigma2) sigma2[l] = ex->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = ex->inv_sigma2[l];
        if (level_w) level_w[l] = ex->G.lv[l].w;
        if (level_h) level_h[l] = ex->G.lv[l].h;
    }
    return ORBX_OK;
}

int orbx_extractor_set_profiling(orbx_extractor *ex, int enabled) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    ex->profiling = enabled != 0; ex->have_times = false; ex->prof_calls = 0;
    if (enabled && ex->d_fast_stats) {
        ORBX_CUDA(cudaSetDevice(ex->device));
        ORBX_CUDA(cudaMemsetAsync(ex->d_fast_stats, 0, sizeof(FastStats), ex->stream));
    }
    return ORBX_OK;
}

// FAST counters accumulated over the profiled calls since profiling was switched on: tested pixels, quick-test survivors, corners
// (score >= threshold), kept after non-max suppression, retried cells (minThFAST), tiles (CTAs).
int orbx_extractor_fast_stats(orbx_extractor *ex, unsigned long long *out6) {
    if (!ex || !out6) return fail(ORBX_ERR_INVALID, "NULL argument");
    ORBX_CUDA(cudaSetDevice(ex->device));
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    ORBX_CUDA(cudaMemcpy(out6, ex->d_fast_stats, sizeof(FastStats), cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_extractor_last_launches(const orbx_extractor *ex) { return ex ? ex->last_launches : 0; }

// Launches of one chunk: images [i0, i0 + cnt) of the batch, pixels at d_img (device, tightly packed, image i0 first).
// sA carries pyramid -> FAST -> quadtree -> slots -> orientation + descriptor; the blur runs on sB between the pyramid and
// the descriptor.  With sA == sB (profiling) the stages are serial and `ev` receives one event per stage boundary.
static int enqueue_chunk(orbx_extractor *ex, int i0, int cnt, const uint8_t *d_img, cudaStream_t sA, cudaStream_t sB, int lane,
                         cudaEvent_t *ev) {
    const Geom &G = ex->G;
    const int nl = ex->nlevels;
    int launches = 0;
    {   // pyramid
        const LevelGeom &L0 = G.lv[0];
        const int chunks = (kXO + L0.w + kEdge + 15) >> 4;
        dim3 grid((chunks * ((L0.rows + kL0Rows - 1) / kL0Rows) + 255) / 256, cnt);
        k_level0<<<grid, 256, 0, sA>>>(d_img, (long long)ex->width * ex->height, ex->width, ex->d_pyr, L0, i0);
        launches++;
        for (int l = 1; l < nl; l++) {
            const LevelGeom &D = G.lv[l];
            dim3 g((D.w + kRszTW - 1) / kRszTW, (D.h + kRszTH - 1) / kRszTH, cnt);
            k_resize<<<g, kRszThreads, ex->rsz_smem, sA>>>(ex->d_pyr, G.lv[l - 1], D, ex->d_taps, ex->rsz_rows, ex->rsz_pitch, ex->rsz_maps[l], i0);
            launches++;
        }
    }
    if (ev) ORBX_CUDA(cudaEventRecord(ev[1], sA));
    // The blur only needs the pyramid.  With overlapping lanes it runs on the aux stream NEXT TO THE QUADTREE (launched after FAST):
    // the quadtree is latency-bound (about a fifth of the issue slots busy), the blur is issue-bound, so they share the SMs well;
    // FAST and the blur would only compete for the same issue slots.
    const bool overlap = sB != sA;
    if (!overlap) {
        k_blur<<<dim3(ex->blur_total_tiles, cnt), kBlurWarps * 32, kBlurSmem, sA>>>(ex->d_pyr, ex->d_blur, G, ex->d_blur_tiles, ex->blur_maps, i0);
        launches++;
    }
    if (ev) ORBX_CUDA(cudaEventRecord(ev[2], sA));
    const FastArgs fa = {G.fast_ymax, G.total_cells, i0, ex->fast_kb_words, G.cand_per_image};
    fast_kernel(G.fast_ip)<<<dim3(G.fast_total_tiles, cnt), kFastThreads, ex->fast_smem, sA>>>(ex->d_pyr, ex->d_fast_tiles, fa, ex->d_params, ex->d_cells, ex->d_cell_cnt, ex->fast_maps,
                                                                                ex->profiling ? ex->d_fast_stats : nullptr);
    launches++;
    if (overlap) {
        ORBX_CUDA(cudaEventRecord(ex->ev_pyr[lane], sA));                 // (FAST done)
        ORBX_CUDA(cudaStreamWaitEvent(sB, ex->ev_pyr[lane], 0));
        k_blur<<<dim3(ex->blur_total_tiles, cnt), kBlurWarps * 32, kBlurSmem, sB>>>(ex->d_pyr, ex->d_blur, G, ex->d_blur_tiles, ex->blur_maps, i0);
        launches++;
        ORBX_CUDA(cudaEventRecord(ex->ev_blur[lane], sB));
    }
    if (ev) ORBX_CUDA(cudaEventRecord(ev[3], sA));
    if (cnt * nl <= ex->oct_wide_max)   // few (image, level) pairs: one wide CTA each (latency), else the throughput shape
        k_octree_wide<<<dim3(cnt, nl), 1024, ex->oct_smem, sA>>>(G, ex->d_nfeat, ex->d_cells, ex->d_cell_cnt, ex->d_cell_off, ex->d_lin,
                                                                 ex->d_lin_n, ex->d_node, ex->d_slot, ex->d_lvl_kp, ex->d_lvl_n,
                                                                 ex->oct_list_cap, i0);
    else
        k_octree<<<dim3(cnt, nl), ex->oct_threads, ex->oct_smem, sA>>>(G, ex->d_nfeat, ex->d_cells, ex->d_cell_cnt, ex->d_cell_off, ex->d_lin,
                                                                       ex->d_lin_n, ex->d_node, ex->d_slot, ex->d_lvl_kp, ex->d_lvl_n,
                                                                       ex->oct_list_cap, i0);
    launches++;
    if (ev) ORBX_CUDA(cudaEventRecord(ev[4], sA));
    k_slots<<<cnt, 256, 0, sA>>>(G, ex->d_params, ex->d_lvl_kp, ex->d_lvl_n, ex->d_slot, ex->d_n_kp, ex->d_n_mono, i0);
    launches++;
    if (ev) ORBX_CUDA(cudaEventRecord(ev[5], sA));
    if (sB != sA) ORBX_CUDA(cudaStreamWaitEvent(sA, ex->ev_blur[lane], 0));
    if (ex->patch_maps.use_tma && ex->desc_batch) {   // one warp per group of kOdG keypoints
        const dim3 grid(ex->desc_blocks2, cnt);
#define ORBX_LAUNCH_DESC_BATCH(FMA, RING)                                                                                                    \
    k_orient_desc_batch<FMA, RING><<<grid, kOdWarps * 32, od_smem_bytes(RING), sA>>>(G, ex->d_pattern, ex->d_lvl_kp, ex->d_lvl_n, ex->d_slot,     \
                                                                                      ex->d_lvl_angle, ex->d_out_kp, ex->d_out_desc, ex->d_patch_maps, \
                                                                                      ex->d_desc_blocks2, i0)
        if (ex->desc_ring == 3) { if (G.fma_mode) ORBX_LAUNCH_DESC_BATCH(true, 3); else ORBX_LAUNCH_DESC_BATCH(false, 3); }
        else if (ex->desc_ring == 4) { if (G.fma_mode) ORBX_LAUNCH_DESC_BATCH(true, 4); else ORBX_LAUNCH_DESC_BATCH(false, 4); }
        else { if (G.fma_mode) ORBX_LAUNCH_DESC_BATCH(true, 5); else ORBX_LAUNCH_DESC_BATCH(false, 5); }
#undef ORBX_LAUNCH_DESC_BATCH
    } else {
        const dim3 grid(ex->desc_blocks, cnt);
#define ORBX_LAUNCH_DESC(FMA, TMA)                                                                                             \
    k_orient_desc<FMA, TMA><<<grid, 256, 0, sA>>>(ex->d_pyr, ex->d_blur, G, ex->d_pattern, ex->d_lvl_kp, ex->d_lvl_n, ex->d_slot, \
                                                  ex->d_lvl_angle, ex->d_out_kp, ex->d_out_desc, ex->d_patch_maps, ex->d_desc_blocks, i0)
        const bool tma = ex->patch_maps.use_tma != 0;
        if (G.fma_mode) { if (tma) ORBX_LAUNCH_DESC(true, true); else ORBX_LAUNCH_DESC(true, false); }
        else { if (tma) ORBX_LAUNCH_DESC(false, true); else ORBX_LAUNCH_DESC(false, false); }
#undef ORBX_LAUNCH_DESC
    }
    launches++;
    if (ev) ORBX_CUDA(cudaEventRecord(ev[6], sA));
    ex->last_launches += launches;
    return ORBX_OK;
}

// enqueue_chunk through a cached CUDA graph for chunks that leave most of the GPU idle (their launches, not their kernels, set the
// pace).  Everything a launch captures by value is constant for the handle (geometry, arenas, tensor maps) or part of the key.
static int enqueue_chunk_graphed(orbx_extractor *ex, int i0, int cnt, const uint8_t *d_img, cudaStream_t sA, cudaStream_t sB, int lane) {
    if (!ex->use_graphs || sA == sB || cnt * ex->nlevels > ex->oct_wide_max) return enqueue_chunk(ex, i0, cnt, d_img, sA, sB, lane, nullptr);
    if (ex->graphs.size() >= 64 && !ex->graphs.count(std::make_tuple(lane, i0, cnt, (const void *)d_img))) {   // callers that keep changing buffers
        for (auto &g : ex->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        ex->graphs.clear();
    }
    orbx_extractor::ChunkGraph &cg = ex->graphs[std::make_tuple(lane, i0, cnt, (const void *)d_img)];
    if (!cg.exec) {
        const int before = ex->last_launches;
        cudaGraph_t graph = nullptr;
        ORBX_CUDA(cudaStreamBeginCapture(sA, cudaStreamCaptureModeThreadLocal));
        const int st = enqueue_chunk(ex, i0, cnt, d_img, sA, sB, lane, nullptr);
        const cudaError_t ce = cudaStreamEndCapture(sA, &graph);
        cg.launches = ex->last_launches - before;
        ex->last_launches = before;
        if (st != ORBX_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ex->graphs.erase(std::make_tuple(lane, i0, cnt, (const void *)d_img));
            ex->use_graphs = false;                                  // fall back to plain launches for this handle
            return enqueue_chunk(ex, i0, cnt, d_img, sA, sB, lane, nullptr);
        }
        const cudaError_t ie = cudaGraphInstantiate(&cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            cudaGetLastError();
            ex->graphs.erase(std::make_tuple(lane, i0, cnt, (const void *)d_img));
            ex->use_graphs = false;
            return enqueue_chunk(ex, i0, cnt, d_img, sA, sB, lane, nullptr);
        }
    }
    ORBX_CUDA(cudaGraphLaunch(cg.exec, sA));
    ex->last_launches += cg.launches;
    return ORBX_OK;
}

static int images_per_chunk(const orbx_extractor *ex, int n, bool host_io) {
    if (ex->chunk_images > 0) return std::min(n, ex->chunk_images);
    if (n <= 16) return n;
    const int parts = host_io ? 4 : 1;     // measured on B200 (320 images of 720x540): 80-image chunks for host I/O, one chunk otherwise
    return std::max(16, (n + parts - 1) / parts);
}

// Validates and stages the per-image parameters (pinned staging -> device, on the handle's stream).
static int stage_params(orbx_extractor *ex, int n, const orbx_image_params *params) {
    if (n <= 0 || n > ex->max_images) return fail(ORBX_ERR_CAPACITY, "n_images exceeds the handle's max_images");
    const Geom &G = ex->G;
    const int nl = ex->nlevels;
    ORBX_CUDA(cudaSetDevice(ex->device));
    // pinned staging of the per-image parameters may still be in flight from the previous call
    ORBX_CUDA(cudaEventSynchronize(ex->ev_staging));
    for (int i = 0; i < n; i++) {
        const orbx_image_params &p = params[i];
        if (p.nfeatures <= 0 || p.nfeatures > ex->max_nfeatures)
            return fail(ORBX_ERR_INVALID, "params[i].nfeatures outside (0, max_nfeatures]");
        ex->h_params[i] = p;
        features_per_level(p.nfeatures, ex->scale_factor, nl, ex->h_nfeat + (size_t)i * nl);
        for (int l = 0; l < nl; l++)
            if (oct_list_cap(ex->h_nfeat[(size_t)i * nl + l], G.lv[l].n_ini) > G.lv[l].kp_cap)
                return fail(ORBX_ERR_CAPACITY, "per-level feature target exceeds the handle's capacity");
    }
    cudaStream_t s = ex->stream;
    ORBX_CUDA(cudaMemcpyAsync(ex->d_params, ex->h_params, n * sizeof(orbx_image_params), cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(ex->d_nfeat, ex->h_nfeat, (size_t)n * nl * sizeof(int), cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaEventRecord(ex->ev_staging, s));
    ex->last_launches = 0;
    ex->last_n = n;
    return ORBX_OK;
}

// Host images of one chunk -> d_in on stream s: one copy per run of images that are contiguous in host memory.
static int upload_chunk(orbx_extractor *ex, int i0, int cnt, const uint8_t *const *images, const int *strides, cudaStream_t s) {
    const size_t img_bytes = (size_t)ex->width * ex->height;
    int i = i0;
    while (i < i0 + cnt) {
        const int st = strides ? strides[i] : ex->width;
        if (st == ex->width) {
            int j = i + 1;
            while (j < i0 + cnt && (strides ? strides[j] : ex->width) == ex->width && images[j] == images[j - 1] + img_bytes) j++;
            ORBX_CUDA(cudaMemcpyAsync(ex->d_in + i * img_bytes, images[i], (size_t)(j - i) * img_bytes, cudaMemcpyHostToDevice, s));
            i = j;
        } else {
            ORBX_CUDA(cudaMemcpy2DAsync(ex->d_in + i * img_bytes, ex->width, images[i], st, ex->width, ex->height, cudaMemcpyHostToDevice, s));
            i++;
        }
    }
    return ORBX_OK;
}

// Results of images [i0, i0 + cnt) -> host buffers on stream s.
static int download_chunk(orbx_extractor *ex, int i0, int cnt, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono, cudaStream_t s) {
    const size_t oc = (size_t)ex->G.out_cap;
    if (kps) {
        if ((size_t)cap == oc)
            ORBX_CUDA(cudaMemcpyAsync(kps + i0 * oc, ex->d_out_kp + i0 * oc, cnt * oc * sizeof(orbx_kp), cudaMemcpyDeviceToHost, s));
        else
            ORBX_CUDA(cudaMemcpy2DAsync(kps + (size_t)i0 * cap, (size_t)cap * sizeof(orbx_kp), ex->d_out_kp + i0 * oc, oc * sizeof(orbx_kp),
                                        oc * sizeof(orbx_kp), cnt, cudaMemcpyDeviceToHost, s));
    }
    if (desc) {
        if ((size_t)cap == oc)
            ORBX_CUDA(cudaMemcpyAsync(desc + i0 * oc * 32, ex->d_out_desc + i0 * oc * 32, cnt * oc * 32, cudaMemcpyDeviceToHost, s));
        else
            ORBX_CUDA(cudaMemcpy2DAsync(desc + (size_t)i0 * cap * 32, (size_t)cap * 32, ex->d_out_desc + i0 * oc * 32, oc * 32, oc * 32, cnt,
                                        cudaMemcpyDeviceToHost, s));
    }
    if (n_kp) ORBX_CUDA(cudaMemcpyAsync(n_kp + i0, ex->d_n_kp + i0, cnt * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (n_mono) ORBX_CUDA(cudaMemcpyAsync(n_mono + i0, ex->d_n_mono + i0, cnt * sizeof(int), cudaMemcpyDeviceToHost, s));
    return ORBX_OK;
}

// The whole batch.  d_images: device pixels (tightly packed) or, when `images` is given, host images that are uploaded
// chunk by chunk into d_in; with host outputs (kps/desc/...) the results of a chunk are downloaded as soon as it is done.
static int run_batch(orbx_extractor *ex, int n, const uint8_t *d_images, const uint8_t *const *images, const int *strides,
                     const orbx_image_params *params, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono) {
    int st = stage_params(ex, n, params);
    if (st != ORBX_OK) return st;
    cudaStream_t s = ex->stream;
    const bool host_io = images != nullptr;
    const size_t img_bytes = (size_t)ex->width * ex->height;
    if (ex->profiling) {   // serial, one chunk, stage events on the handle's stream
        cudaEvent_t *ev = ex->ev[ex->prof_calls % orbx_extractor::kRing];
        if (host_io && (st = upload_chunk(ex, 0, n, images, strides, s)) != ORBX_OK) return st;
        ORBX_CUDA(cudaEventRecord(ev[0], s));
        if ((st = enqueue_chunk(ex, 0, n, host_io ? ex->d_in : d_images, s, s, 0, ev)) != ORBX_OK) return st;
        ex->have_times = true; ex->prof_calls++;
        if (host_io && (st = download_chunk(ex, 0, n, kps, desc, cap, n_kp, n_mono, s)) != ORBX_OK) return st;
        ORBX_CUDA(cudaGetLastError());
        return ORBX_OK;
    }
    const int per = images_per_chunk(ex, n, host_io);
    const int nchunks = (n + per - 1) / per;
    while ((int)ex->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        ORBX_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ORBX_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ex->ev_in.push_back(a); ex->ev_out.push_back(b);
    }
    ORBX_CUDA(cudaEventRecord(ex->ev_fork, s));
    if (host_io) {
        ORBX_CUDA(cudaStreamWaitEvent(ex->s_h2d, ex->ev_fork, 0));
        ORBX_CUDA(cudaStreamWaitEvent(ex->s_d2h, ex->ev_fork, 0));
    }
    const int lanes = std::min(nchunks, (int)orbx_extractor::kLanes);
    for (int l = 0; l < lanes; l++) ORBX_CUDA(cudaStreamWaitEvent(ex->lane_main[l], ex->ev_fork, 0));
    for (int k = 0; k < nchunks; k++) {
        const int i0 = k * per, cnt = std::min(per, n - i0), lane = k % orbx_extractor::kLanes;
        cudaStream_t sA = ex->lane_main[lane], sB = ex->lane_aux[lane];
        if (host_io) {
            if ((st = upload_chunk(ex, i0, cnt, images, strides, ex->s_h2d)) != ORBX_OK) return st;
            ORBX_CUDA(cudaEventRecord(ex->ev_in[k], ex->s_h2d));
            ORBX_CUDA(cudaStreamWaitEvent(sA, ex->ev_in[k], 0));
        }
        if ((st = enqueue_chunk_graphed(ex, i0, cnt, (host_io ? ex->d_in : d_images) + i0 * img_bytes, sA, sB, lane)) != ORBX_OK) return st;
        if (host_io) {
            ORBX_CUDA(cudaEventRecord(ex->ev_out[k], sA));
            ORBX_CUDA(cudaStreamWaitEvent(ex->s_d2h, ex->ev_out[k], 0));
            if ((st = download_chunk(ex, i0, cnt, kps, desc, cap, n_kp, n_mono, ex->s_d2h)) != ORBX_OK) return st;
        }
    }
    for (int l = 0; l < lanes; l++) {   // join: everything is ordered before later work on the handle's stream
        ORBX_CUDA(cudaEventRecord(ex->ev_lane[l], ex->lane_main[l]));
        ORBX_CUDA(cudaStreamWaitEvent(s, ex->ev_lane[l], 0));
    }
    if (host_io) {
        ORBX_CUDA(cudaEventRecord(ex->ev_d2h, ex->s_d2h));
        ORBX_CUDA(cudaStreamWaitEvent(s, ex->ev_d2h, 0));
    }
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

int orbx_extract_batch_device(orbx_extractor *ex, int n_images, const uint8_t *d_images, const orbx_image_params *params) {
    if (!ex || !d_images || !params) return fail(ORBX_ERR_INVALID, "NULL argument");
    return run_batch(ex, n_images, d_images, nullptr, nullptr, params, nullptr, nullptr, 0, nullptr, nullptr);
}

int orbx_extract_fetch(orbx_extractor *ex, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    const int n = ex->last_n;
    if (n <= 0) return fail(ORBX_ERR_INVALID, "no results to fetch");
    if (cap < ex->G.out_cap && (kps || desc)) return fail(ORBX_ERR_CAPACITY, "cap smaller than orbx_extractor_capacity()");
    ORBX_CUDA(cudaSetDevice(ex->device));
    int st = download_chunk(ex, 0, n, kps, desc, cap, n_kp, n_mono, ex->stream);
    if (st != ORBX_OK) return st;
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBX_OK;
}

int orbx_extract_device_results(orbx_extractor *ex, const orbx_kp **d_kps, const uint8_t **d_desc, const int **d_n_kp,
                                const int **d_n_mono) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    if (d_kps) *d_kps = ex->d_out_kp;
    if (d_desc) *d_desc = ex->d_out_desc;
    if (d_n_kp) *d_n_kp = ex->d_n_kp;
    if (d_n_mono) *d_n_mono = ex->d_n_mono;
    return ORBX_OK;
}

int orbx_extract_batch_submit(orbx_extractor *ex, int n_images, const uint8_t *const *images, const int *strides,
                              const orbx_image_params *params, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono) {
    if (!ex || !images || !params) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (n_images <= 0 || n_images > ex->max_images) return fail(ORBX_ERR_CAPACITY, "n_images exceeds the handle's max_images");
    if (cap < ex->G.out_cap && (kps || desc)) return fail(ORBX_ERR_CAPACITY, "cap smaller than orbx_extractor_capacity()");
    for (int i = 0; i < n_images; i++) {
        if (!images[i]) return fail(ORBX_ERR_EMPTY, "empty image");
        if (strides && strides[i] < ex->width) return fail(ORBX_ERR_INVALID, "stride smaller than width");
    }
    ORBX_CUDA(cudaSetDevice(ex->device));
    return run_batch(ex, n_images, nullptr, images, strides, params, kps, desc, cap, n_kp, n_mono);
}

int orbx_extract_wait(orbx_extractor *ex) {
    if (!ex) return fail(ORBX_ERR_INVALID, "NULL handle");
    ORBX_CUDA(cudaSetDevice(ex->device));
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBX_OK;
}

int orbx_extract_batch(orbx_extractor *ex, int n_images, const uint8_t *const *images, const int *strides,
                       const orbx_image_params *params, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono) {
    int st = orbx_extract_batch_submit(ex, n_images, images, strides, params, kps, desc, cap, n_kp, n_mono);
    if (st != ORBX_OK) return st;
    return orbx_extract_wait(ex);
}

int orbx_extract(orbx_extractor *ex, const uint8_t *image, int stride, const orbx_image_params *params, orbx_kp *kps,
                 uint8_t *desc, int cap, int *n_kp) {
    if (!ex || !params) { fail(ORBX_ERR_INVALID, "NULL handle or parameters"); return ORBX_ERR_ARGUMENT; }
    if (!image) return -1;   // src/ORBextractor.cc:994-995
    int mono = 0, nk = 0;
    const uint8_t *imgs[1] = {image};
    int st = orbx_extract_batch(ex, 1, imgs, &stride, params, kps, desc, cap, &nk, &mono);
    if (st == ORBX_ERR_INVALID) return ORBX_ERR_ARGUMENT;     // -1 is taken by the empty-image return value above
    if (st != ORBX_OK) return st;
    if (n_kp) *n_kp = nk;
    return mono;
}

int orbx_extractor_pyramid_level(orbx_extractor *ex, int image, int level, int with_border, uint8_t *dst, int dst_stride) {
    if (!ex || !dst) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (image < 0 || image >= ex->max_images || level < 0 || level >= ex->nlevels) return fail(ORBX_ERR_INVALID, "bad index");
    const LevelGeom &L = ex->G.lv[level];
    ORBX_CUDA(cudaSetDevice(ex->device));
    const uint8_t *p = ex->d_pyr + L.plane_off + (long long)image * L.plane_bytes;
    const int w = with_border ? L.w + 2 * kEdge : L.w, h = with_border ? L.h + 2 * kEdge : L.h;
    if (dst_stride < w) return fail(ORBX_ERR_INVALID, "dst_stride too small");
    const uint8_t *src = with_border ? p + (kXO - kEdge) : p + (long long)kYO * L.pitch + kXO;
    ORBX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, src, L.pitch, w, h, cudaMemcpyDeviceToHost, ex->stream));
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBX_OK;
}

int orbx_debug_blur_level(orbx_extractor *ex, int image, int level, uint8_t *dst, int dst_stride) {
    if (!ex || !dst) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (image < 0 || image >= ex->max_images || level < 0 || level >= ex->nlevels) return fail(ORBX_ERR_INVALID, "bad index");
    const LevelGeom &L = ex->G.lv[level];
    ORBX_CUDA(cudaSetDevice(ex->device));
    const uint8_t *p = ex->d_blur + L.plane_off + (long long)image * L.plane_bytes + (long long)kYO * L.pitch + kXO;
    ORBX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, p, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, ex->stream));
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    return ORBX_OK;
}

int orbx_debug_candidates(orbx_extractor *ex, int image, int level, int *xs, int *ys, int *scores, int cap, int *n) {
    if (!ex || !n) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (image < 0 || image >= ex->max_images || level < 0 || level >= ex->nlevels) return fail(ORBX_ERR_INVALID, "bad index");
    const LevelGeom &L = ex->G.lv[level];
    ORBX_CUDA(cudaSetDevice(ex->device));
    int cnt = 0;
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    ORBX_CUDA(cudaMemcpy(&cnt, ex->d_lin_n + image * ex->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    const int m = std::min(cnt, cap);
    if (m > 0 && xs && ys && scores) {
        std::vector<uint32_t> v(m);
        ORBX_CUDA(cudaMemcpy(v.data(), ex->d_lin + (long long)image * ex->G.cand_per_image + L.cand_off, m * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) { xs[i] = cand_x(v[i]); ys[i] = cand_y(v[i]); scores[i] = cand_score(v[i]); }
    }
    return ORBX_OK;
}

int orbx_debug_level_keypoints(orbx_extractor *ex, int image, int level, orbx_kp *kps, int cap, int *n) {
    if (!ex || !n) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (image < 0 || image >= ex->max_images || level < 0 || level >= ex->nlevels) return fail(ORBX_ERR_INVALID, "bad index");
    const LevelGeom &L = ex->G.lv[level];
    ORBX_CUDA(cudaSetDevice(ex->device));
    ORBX_CUDA(cudaStreamSynchronize(ex->stream));
    int cnt = 0;
    ORBX_CUDA(cudaMemcpy(&cnt, ex->d_lvl_n + image * ex->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    const int m = std::min(cnt, cap);
    if (m > 0 && kps) {
        std::vector<uint32_t> v(m);
        std::vector<float> a(m);
        const long long off = (long long)image * ex->G.kp_per_image + L.kp_off;
        ORBX_CUDA(cudaMemcpy(v.data(), ex->d_lvl_kp + off, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        ORBX_CUDA(cudaMemcpy(a.data(), ex->d_lvl_angle + off, m * sizeof(float), cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) {
            kps[i].x = (float)cand_x(v[i]); kps[i].y = (float)cand_y(v[i]); kps[i].size = L.patch; kps[i].angle = a[i];
            kps[i].response = (float)cand_score(v[i]); kps[i].octave = level; kps[i].class_id = -1;
        }
    }
    return ORBX_OK;
}

int orbx_compute_stereo_matches2(orbx_extractor *exl, int img_left, orbx_extractor *exr, int img_right, float mb, float mbf,
                                 float *u_right, float *depth, int cap, int *n_matches) {
    if (!exl || !exr || !u_right || !depth) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (img_left < 0 || img_right < 0 || img_left >= exl->last_n || img_right >= exr->last_n)
        return fail(ORBX_ERR_INVALID, "image index outside the last batch");
    if (exl->device != exr->device || exl->width != exr->width || exl->height != exr->height || exl->nlevels != exr->nlevels ||
        exl->scale_factor != exr->scale_factor)
        return fail(ORBX_ERR_INVALID, "left and right extractor differ in device, image size or pyramid shape");
    if (cap < exl->G.out_cap) return fail(ORBX_ERR_CAPACITY, "cap smaller than orbx_extractor_capacity()");
    if (!(mb > 0.0f) || exl->G.out_cap >= 65535 || exr->G.out_cap >= 65535) return fail(ORBX_ERR_INVALID, "bad baseline / too many keypoints");
    ORBX_CUDA(cudaSetDevice(exl->device));
    if (exr != exl) ORBX_CUDA(cudaStreamSynchronize(exr->stream));   // the right extraction must be complete
    const size_t oc = (size_t)exl->G.out_cap, ocr = (size_t)exr->G.out_cap;
    if (!exl->d_st_uright) {
        ORBX_CUDA(dalloc(&exl->d_st_uright, oc)); ORBX_CUDA(dalloc(&exl->d_st_depth, oc));
        ORBX_CUDA(dalloc(&exl->d_st_sad, oc)); ORBX_CUDA(dalloc(&exl->d_st_kept, (size_t)1));
    }
    StereoArgs A;
    A.kpL = exl->d_out_kp + img_left * oc; A.kpR = exr->d_out_kp + img_right * ocr;
    A.descL = reinterpret_cast<const uint4 *>(exl->d_out_desc + img_left * oc * 32);
    A.descR = reinterpret_cast<const uint4 *>(exr->d_out_desc + img_right * ocr * 32);
    A.nL = exl->d_n_kp + img_left; A.nR = exr->d_n_kp + img_right;
    A.pyrL = exl->d_pyr; A.pyrR = exr->d_pyr;
    A.imgL = img_left; A.imgR = img_right;
    for (int l = 0; l < kMaxLevels; l++) A.inv_sf[l] = l < exl->nlevels ? exl->inv_sf[l] : 0.f;
    A.maxD = mbf / mb; A.mbf = mbf;                                  // src/Frame.cc:1086-1088
    A.uRight = exl->d_st_uright; A.depth = exl->d_st_depth; A.sad = exl->d_st_sad;
    cudaStream_t s = exl->stream;
    // the plane geometry of both arenas is the same only if both handles were created for the same max_images
    Geom G = exl->G;
    if (exr != exl && exr->max_images != exl->max_images) return fail(ORBX_ERR_INVALID, "left and right extractor differ in max_images");
    k_stereo_match<<<(unsigned)((oc * 32 + 255) / 256), 256, 0, s>>>(G, A);
    k_stereo_filter<<<1, 1024, 0, s>>>(A, exl->d_st_kept);
    ORBX_CUDA(cudaGetLastError());
    ORBX_CUDA(cudaMemcpyAsync(u_right, exl->d_st_uright, oc * sizeof(float), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(depth, exl->d_st_depth, oc * sizeof(float), cudaMemcpyDeviceToHost, s));
    int kept = 0;
    ORBX_CUDA(cudaMemcpyAsync(&kept, exl->d_st_kept, sizeof(int), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    if (n_matches) *n_matches = kept;
    return ORBX_OK;
}

int orbx_compute_stereo_matches(orbx_extractor *ex, int img_left, int img_right, float mb, float mbf, float *u_right, float *depth,
                                int cap, int *n_matches) {
    return orbx_compute_stereo_matches2(ex, img_left, ex, img_right, mb, mbf, u_right, depth, cap, n_matches);
}

int orbx_extractor_stage_ms(orbx_extractor *ex, float *ms) {
    if (!ex || !ms) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (!ex->have_times) return fail(ORBX_ERR_INVALID, "profiling was not enabled for the last call");
    ORBX_CUDA(cudaSetDevice(ex->device));
    // average over the profiled calls still in the ring (at most kRing most recent ones)
    const int have = std::min(ex->prof_calls, (int)orbx_extractor::kRing);
    for (int i = 0; i < ORBX_NUM_STAGES; i++) ms[i] = 0.f;
    for (int c = 0; c < have; c++) {
        cudaEvent_t *ev = ex->ev[(ex->prof_calls - 1 - c) % orbx_extractor::kRing];
        ORBX_CUDA(cudaEventSynchronize(ev[6]));
        for (int i = 0; i < 6; i++) { float t = 0; ORBX_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1])); ms[i] += t / have; }
        float t = 0; ORBX_CUDA(cudaEventElapsedTime(&t, ev[0], ev[6])); ms[6] += t / have;
    }
    return ORBX_OK;
}

}  // extern "C"
