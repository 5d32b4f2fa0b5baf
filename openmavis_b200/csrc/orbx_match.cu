// orbx_match.cu — host side of the matching C ABI (include/orbx.h).
// Replaces the Hamming cores of ORBmatcher (src/ORBmatcher.cc) and cv::BFMatcher::knnMatch (src/Frame.cc:1427,1483).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "orbx_match_kernels.cuh"
#include "orbx_sbp_kernels.cuh"
#include "orbx_fisheye_kernels.cuh"
#include "orbx_bow_kernels.cuh"
#include "orbx_sft_kernels.cuh"

using namespace orbx;

namespace {
// growable device buffer
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max<size_t>(bytes + bytes / 4, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() { return (T *)p; }
};
}  // namespace

struct orbx_vocabulary {
    int device = 0, n_nodes = 0, n_children = 0, levels = 0;
    int *d_child_ptr = nullptr, *d_child = nullptr, *d_word = nullptr;
    uint8_t *d_desc = nullptr;
    std::vector<double> weight;      // per node (leaves); looked up on the host
    std::vector<int> word;
    std::vector<int> leaf_of_word;   // word id -> leaf node (m_words order), built once in orbx_vocabulary_create
};

struct orbx_matcher {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr, aux_stream = nullptr;   // aux: the peer pull of the sharded 2-NN
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    int last_launches = 0;
    int knn_resident = 8;                           // CTAs of k_knn2 one SM holds
    bool timed = false;
    DBuf b[24];
    std::vector<uint8_t> last_cleared;              // orbx_search_by_projection_last: keypoints whose claim the rotation filter cleared
    uint8_t *stage_h = nullptr;                     // pinned staging of the small matchers' inputs: ONE host-to-device copy per call (Packer)
    size_t stage_hcap = 0;
    DBuf stage_d;
};

#define M_CUDA ORBX_CUDA

static int upload(orbx_matcher *m, DBuf &buf, const void *src, size_t bytes) {
    M_CUDA(buf.ensure(bytes));
    if (bytes) M_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, m->stream));
    return ORBX_OK;
}

// The guided matchers of one frame are a dozen small arrays in, a few kernels, one array out: a dozen cudaMemcpyAsync from pageable
// memory cost more than the kernels (each is staged by the driver on its own).  A Packer gathers the arrays of a call in the handle's
// pinned staging buffer (256-byte aligned offsets) and moves them with ONE asynchronous copy into one device arena; the call's final
// stream synchronisation (every entry point ends with one) makes the staging buffer free again.
struct Packer {
    struct Item { size_t off; const void *src; size_t bytes; };
    std::vector<Item> items;
    size_t total = 0;
    size_t add(const void *src, size_t bytes) {
        const size_t o = total;
        items.push_back({o, src, bytes});
        total += (bytes + 255) & ~(size_t)255;
        return o;
    }
    int commit(orbx_matcher *m, uint8_t **base) {
        const size_t need = std::max<size_t>(total, 256);
        if (need > m->stage_hcap) {
            if (m->stage_h) cudaFreeHost(m->stage_h);
            m->stage_h = nullptr; m->stage_hcap = 0;
            const size_t want = need + need / 4;
            M_CUDA(cudaHostAlloc((void **)&m->stage_h, want, cudaHostAllocDefault));
            m->stage_hcap = want;
        }
        M_CUDA(m->stage_d.ensure(need));
        for (const Item &it : items)
            if (it.bytes) memcpy(m->stage_h + it.off, it.src, it.bytes);
        if (total) M_CUDA(cudaMemcpyAsync(m->stage_d.p, m->stage_h, total, cudaMemcpyHostToDevice, m->stream));
        *base = m->stage_d.as<uint8_t>();
        return ORBX_OK;
    }
};

// Train set = concatenation of `nparts` parts (one for the single-GPU call; one per rank for the sharded matcher).  The parts are
// matched in at most two launches — the part `first` alone (or nothing: -1), then all the others in ONE launch after the stream has
// waited for `wait_after_first` (the peer rows arrive on another stream while the own part is being matched) — every launch cut into
// enough train slices for two full waves of resident CTAs, all slices writing one partial buffer, one merge.
static int knn2_launch_parts(orbx_matcher *m, const uint8_t *d_q, int nq, const uint8_t *const *d_parts, const int *part_rows, int nparts,
                             int *d_idx, int *d_dist, int first = -1, cudaEvent_t wait_after_first = nullptr) {
    long long nt = 0;
    for (int p = 0; p < nparts; p++) nt += part_rows[p];
    if (nt >= (1 << kKeyShift)) return fail(ORBX_ERR_INVALID, "knn2: more than 8388607 train rows are not supported");
    const int per_block = kKnnThreads * kKnnQPT;
    const int qblocks = (nq + per_block - 1) / per_block;
    std::vector<int> row_off(nparts);
    for (int p = 0, ro = 0; p < nparts; p++) { row_off[p] = ro; ro += part_rows[p]; }
    // the two groups and their slices
    KnnGroup G[2];
    int total_slices = 0;
    for (int g = 0; g < 2; g++) {
        KnnGroup &K = G[g];
        K.nparts = 0; K.first_slice[0] = total_slices;
        long long rows_g = 0;
        for (int p = 0; p < nparts; p++) if ((p == first) == (g == 0) && part_rows[p] > 0) rows_g += part_rows[p];
        if (rows_g == 0) continue;
        // slices of the group: two waves of the CTAs the device holds at once, at least one tile per slice
        const int want = std::max(1, (2 * m->sm_count * m->knn_resident + qblocks - 1) / qblocks);
        for (int p = 0; p < nparts; p++) {
            if ((p == first) != (g == 0) || part_rows[p] <= 0) continue;
            const int rows = part_rows[p], tiles = (rows + kKnnTile - 1) / kKnnTile;
            const int s = std::max(1, std::min((int)((long long)want * rows / rows_g) + 1, tiles));
            const int srows = ((rows + s - 1) / s + kKnnTile - 1) / kKnnTile * kKnnTile;
            const int k = K.nparts++;
            K.t[k] = (const uint4 *)d_parts[p]; K.nt[k] = rows; K.row_off[k] = row_off[p]; K.slice_rows[k] = srows;
            total_slices += (rows + srows - 1) / srows;
            K.first_slice[k + 1] = total_slices;
        }
    }
    M_CUDA(m->b[0].ensure((size_t)std::max(total_slices, 1) * nq * sizeof(uint2)));
    for (int g = 0; g < 2; g++) {
        if (g == 1 && wait_after_first) M_CUDA(cudaStreamWaitEvent(m->stream, wait_after_first, 0));
        const KnnGroup &K = G[g];
        if (K.nparts == 0) continue;
        const int ns = K.first_slice[K.nparts] - K.first_slice[0];
        KnnGroup L = K;                              // slice numbers relative to the launch
        for (int k = 0; k <= L.nparts; k++) L.first_slice[k] -= K.first_slice[0];
        k_knn2<<<dim3(qblocks, ns), kKnnThreads, 0, m->stream>>>((const uint4 *)d_q, nq, L, m->b[0].as<uint2>() + (size_t)K.first_slice[0] * nq);
        m->last_launches++;
    }
    k_knn2_merge<<<(nq + 255) / 256, 256, 0, m->stream>>>(m->b[0].as<uint2>(), nq, total_slices, d_idx, d_dist);
    m->last_launches++;
    M_CUDA(cudaGetLastError());
    return ORBX_OK;
}

static int knn2_launch(orbx_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, int *d_idx, int *d_dist) {
    return knn2_launch_parts(m, d_q, nq, &d_t, &nt, 1, d_idx, d_dist);
}

extern "C" {

void orbx_matcher_destroy(orbx_matcher *m);
int orbx_matcher_create(orbx_matcher **out, int device) {
    if (!out) return fail(ORBX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    const int ndev = orbx_device_count();
    if (ndev <= 0) return fail(ORBX_ERR_NO_DEVICE, "no CUDA device (liborbx has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ORBX_ERR_NO_DEVICE, "device index out of range");
    M_CUDA(cudaSetDevice(device));
    orbx_matcher *m = new orbx_matcher();
    m->device = device;
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) {
        orbx_matcher_destroy(m);
        return fail(ORBX_ERR_CUDA, "stream/event creation failed");
    }
    m->stream = m->own_stream;
    {   // load the kernels of the sharded 2-NN now: with lazy module loading the FIRST launch of a kernel may have to wait for running
        // kernels, and the pull kernel of that path spins on a flag that another kernel publishes
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_publish_epoch); cudaFuncGetAttributes(&fa, k_pull_parts_wait); cudaFuncGetAttributes(&fa, k_pull_parts);
        cudaFuncGetAttributes(&fa, k_knn2); cudaFuncGetAttributes(&fa, k_knn2_merge);
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_knn2, kKnnThreads, 0) == cudaSuccess && nb > 0) m->knn_resident = nb;
    }
    *out = m;
    return ORBX_OK;
}

void orbx_matcher_destroy(orbx_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (auto &b : m->b) b.release();
    m->stage_d.release();
    if (m->stage_h) cudaFreeHost(m->stage_h);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

int orbx_matcher_set_stream(orbx_matcher *m, void *cuda_stream) {
    if (!m) return fail(ORBX_ERR_INVALID, "NULL handle");
    m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
    return ORBX_OK;
}

int orbx_matcher_last_launches(const orbx_matcher *m) { return m ? m->last_launches : 0; }

int orbx_matcher_last_ms(orbx_matcher *m, float *ms) {
    if (!m || !ms) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (!m->timed) return fail(ORBX_ERR_INVALID, "no timed call yet");
    M_CUDA(cudaEventSynchronize(m->ev1));
    M_CUDA(cudaEventElapsedTime(ms, m->ev0, m->ev1));
    return ORBX_OK;
}

int orbx_descriptor_distance(int device, const uint8_t *a, const uint8_t *b, int n, int *dist) {
    if (!a || !b || !dist || n < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    if (orbx_device_count() <= device || device < 0) return fail(ORBX_ERR_NO_DEVICE, "no such CUDA device");
    if (n == 0) return ORBX_OK;
    M_CUDA(cudaSetDevice(device));
    // one allocation for the three arrays: nothing to leak on an error path
    uint8_t *buf = nullptr;
    const size_t nb = (size_t)n * 32;
    M_CUDA(cudaMalloc((void **)&buf, 2 * nb + (size_t)n * sizeof(int)));
    cudaError_t e = cudaMemcpy(buf, a, nb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(buf + nb, b, nb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_pair_distance<<<(n + 255) / 256, 256>>>((const uint4 *)buf, (const uint4 *)(buf + nb), n, (int *)(buf + 2 * nb));
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(dist, buf + 2 * nb, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    M_CUDA(e);
    return ORBX_OK;
}

int orbx_knn2_device(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_train, int nt, int *d_idx, int *d_dist) {
    if (!m || nq < 0 || nt < 0 || !d_idx || !d_dist) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (nq == 0) return ORBX_OK;
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    if (nt == 0) {
        M_CUDA(cudaMemsetAsync(d_idx, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
        M_CUDA(cudaMemsetAsync(d_dist, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
    } else {
        int st = knn2_launch(m, d_query, nq, d_train, nt, d_idx, d_dist);
        if (st != ORBX_OK) return st;
    }
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    return ORBX_OK;
}

int orbx_knn2_device_parts(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *const *d_train_parts, const int *part_rows,
                           int nparts, int *d_idx, int *d_dist) {
    if (!m || nq < 0 || nparts < 1 || nparts > 64 || !d_train_parts || !part_rows || !d_idx || !d_dist) return fail(ORBX_ERR_INVALID, "bad arguments");
    long long nt = 0;
    for (int p = 0; p < nparts; p++) { if (part_rows[p] < 0 || (part_rows[p] && !d_train_parts[p])) return fail(ORBX_ERR_INVALID, "bad part"); nt += part_rows[p]; }
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (nq == 0) return ORBX_OK;
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    if (nt == 0) {
        M_CUDA(cudaMemsetAsync(d_idx, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
        M_CUDA(cudaMemsetAsync(d_dist, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
    } else if (nparts == 1 || getenv("ORBX_KNN_NO_PULL")) {
        int st = knn2_launch_parts(m, d_query, nq, d_train_parts, part_rows, nparts, d_idx, d_dist);
        if (st != ORBX_OK) return st;
    } else {
        // pull the parts once into a contiguous local buffer (peer loads over NVLink, stream-ordered, no collective call),
        // then match against the local copy: every peer row crosses the link once instead of once per query block
        KnnParts P;
        P.nparts = nparts; P.first_vec[0] = 0;
        for (int p = 0; p < nparts; p++) { P.ptr[p] = (const uint4 *)d_train_parts[p]; P.first_vec[p + 1] = P.first_vec[p] + 2 * part_rows[p]; }
        M_CUDA(m->b[2].ensure((size_t)nt * 32));
        k_pull_parts<<<std::min(m->sm_count * 4, (P.first_vec[nparts] + 255) / 256), 256, 0, m->stream>>>(P, m->b[2].as<uint4>());
        m->last_launches++;
        int st = knn2_launch(m, d_query, nq, m->b[2].as<uint8_t>(), (int)nt, d_idx, d_dist);
        if (st != ORBX_OK) return st;
    }
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    return ORBX_OK;
}

int orbx_knn2_device_parts_sync(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *const *d_train_parts, const int *part_rows,
                                int nparts, int self, const uint32_t *const *d_flags, uint32_t epoch, double timeout_s, int *d_idx, int *d_dist) {
    if (!m || nq < 0 || nparts < 1 || nparts > 64 || !d_train_parts || !part_rows || !d_idx || !d_dist || !d_flags || self < -1 || self >= nparts)
        return fail(ORBX_ERR_INVALID, "bad arguments");
    long long nt = 0;
    for (int p = 0; p < nparts; p++) {
        if (part_rows[p] < 0 || (part_rows[p] && (!d_train_parts[p] || (p != self && !d_flags[p])))) return fail(ORBX_ERR_INVALID, "bad part");
        nt += part_rows[p];
    }
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (nq == 0) return ORBX_OK;
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    if (nt == 0) {
        M_CUDA(cudaMemsetAsync(d_idx, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
        M_CUDA(cudaMemsetAsync(d_dist, 0xFF, (size_t)nq * 2 * sizeof(int), m->stream));
    } else {
        // aux stream: wait for the peers' epoch flags and pull their rows into a local buffer laid out like the whole train set;
        // main stream: match the own part in place meanwhile, then the pulled parts, then one merge
        KnnParts P;
        KnnFlags F;
        P.nparts = nparts; P.first_vec[0] = 0;
        std::vector<const uint8_t *> local(nparts);
        M_CUDA(m->b[2].ensure((size_t)nt * 32));
        for (int p = 0; p < nparts; p++) {
            P.ptr[p] = (const uint4 *)d_train_parts[p]; P.first_vec[p + 1] = P.first_vec[p] + 2 * part_rows[p];
            F.flag[p] = d_flags[p];
            local[p] = p == self ? d_train_parts[p] : m->b[2].as<uint8_t>() + (size_t)P.first_vec[p] * 16;
        }
        F.epoch = epoch; F.self = self;
        F.timeout_ns = (unsigned long long)(std::max(0.001, timeout_s > 0 ? timeout_s : 10.0) * 1e9);
        M_CUDA(cudaEventRecord(m->ev_fork, m->stream));
        M_CUDA(cudaStreamWaitEvent(m->aux_stream, m->ev_fork, 0));
        k_pull_parts_wait<<<std::min(m->sm_count, (P.first_vec[nparts] + 255) / 256), 256, 0, m->aux_stream>>>(P, F, m->b[2].as<uint4>());
        m->last_launches++;
        M_CUDA(cudaEventRecord(m->ev_join, m->aux_stream));
        int st = knn2_launch_parts(m, d_query, nq, local.data(), part_rows, nparts, d_idx, d_dist, self, m->ev_join);
        if (st != ORBX_OK) return st;
    }
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    return ORBX_OK;
}

int orbx_shared_publish(int device, void *d_flag, uint32_t epoch, void *cuda_stream) {
    if (!d_flag) return fail(ORBX_ERR_INVALID, "flag is NULL");
    M_CUDA(cudaSetDevice(device));
    k_publish_epoch<<<1, 1, 0, (cudaStream_t)cuda_stream>>>((uint32_t *)d_flag, epoch);
    M_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// ---- device buffers that peer processes (one per GPU) can map: the train rows of the sharded brute-force matcher are read
// in place over NVLink instead of being all-gathered first
int orbx_shared_alloc(int device, size_t bytes, void **d_ptr, void *ipc_handle64) {
    if (!d_ptr || !ipc_handle64 || bytes == 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    M_CUDA(cudaSetDevice(device));
    M_CUDA(cudaMalloc(d_ptr, bytes));
    M_CUDA(cudaMemset(*d_ptr, 0, bytes));              // (epoch flags start at 0)
    { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_publish_epoch); }   // (loaded before anybody can wait for it)
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return fail(ORBX_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(ipc_handle64, &h, 64);
    return ORBX_OK;
}
int orbx_shared_open(int device, const void *ipc_handle64, void **d_ptr) {
    if (!d_ptr || !ipc_handle64) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    M_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ORBX_OK;
}
int orbx_shared_close(int device, void *d_ptr) {
    M_CUDA(cudaSetDevice(device));
    M_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return ORBX_OK;
}
int orbx_device_copy(int device, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream) {
    M_CUDA(cudaSetDevice(device));
    M_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
    return ORBX_OK;
}
int orbx_shared_free(int device, void *d_ptr) {
    M_CUDA(cudaSetDevice(device));
    M_CUDA(cudaFree(d_ptr));
    return ORBX_OK;
}

int orbx_knn2(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, int *idx, int *dist) {
    if (!m || nq < 0 || nt < 0 || !idx || !dist || (nq && !query) || (nt && !train)) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    if (nq == 0) return ORBX_OK;
    int st;
    if ((st = upload(m, m->b[1], query, (size_t)nq * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], train, (size_t)nt * 32)) != ORBX_OK) return st;
    M_CUDA(m->b[3].ensure((size_t)nq * 2 * sizeof(int)));
    M_CUDA(m->b[4].ensure((size_t)nq * 2 * sizeof(int)));
    st = orbx_knn2_device(m, m->b[1].as<uint8_t>(), nq, m->b[2].as<uint8_t>(), nt, m->b[3].as<int>(), m->b[4].as<int>());
    if (st != ORBX_OK) return st;
    M_CUDA(cudaMemcpyAsync(idx, m->b[3].p, (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(dist, m->b[4].p, (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_ratio_test(const int *dist, int nq, double ratio, uint8_t *keep) {
    if (!dist || !keep || nq < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    for (int i = 0; i < nq; i++)   // (*it)[0].distance < (*it)[1].distance * 0.7 : float * double
        keep[i] = (dist[2 * i] >= 0 && dist[2 * i + 1] >= 0 && (double)(float)dist[2 * i] < (double)(float)dist[2 * i + 1] * ratio) ? 1 : 0;
    return ORBX_OK;
}

static int match_csr_impl(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, const int *row_ptr,
                          const int *cand, int *best_idx, int *best_dist, int *second_dist, int *best_pos, int *second_pos,
                          int *all_dist) {
    if (!m || nq < 0 || nt < 0 || !row_ptr || (nq && !query)) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (nq == 0) return ORBX_OK;
    const int total = row_ptr[nq];
    if (total < 0 || (total && (!cand || !train))) return fail(ORBX_ERR_INVALID, "bad candidate lists");
    for (int k = 0; k < total; k++)
        if (cand[k] < 0 || cand[k] >= nt) return fail(ORBX_ERR_INVALID, "candidate index out of range");
    int st;
    if ((st = upload(m, m->b[1], query, (size_t)nq * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], train, (size_t)nt * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[5], row_ptr, (size_t)(nq + 1) * sizeof(int))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[6], cand, (size_t)total * sizeof(int))) != ORBX_OK) return st;
    M_CUDA(m->b[7].ensure((size_t)nq * 5 * sizeof(int)));
    M_CUDA(m->b[8].ensure((size_t)std::max(total, 1) * sizeof(int)));
    int *o = m->b[7].as<int>();
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_match_csr<<<(nq + 127) / 128, 128, 0, m->stream>>>((const uint4 *)m->b[1].p, nq, (const uint4 *)m->b[2].p, m->b[5].as<int>(),
                                                         m->b[6].as<int>(), best_idx ? o : nullptr, o + nq, o + 2 * nq, o + 3 * nq,
                                                         o + 4 * nq, all_dist ? m->b[8].as<int>() : nullptr);
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    const size_t nb = (size_t)nq * sizeof(int);
    if (best_idx) {
        M_CUDA(cudaMemcpyAsync(best_idx, o, nb, cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaMemcpyAsync(best_dist, o + nq, nb, cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaMemcpyAsync(second_dist, o + 2 * nq, nb, cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaMemcpyAsync(best_pos, o + 3 * nq, nb, cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaMemcpyAsync(second_pos, o + 4 * nq, nb, cudaMemcpyDeviceToHost, m->stream));
    }
    if (all_dist && total) M_CUDA(cudaMemcpyAsync(all_dist, m->b[8].p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_match_csr(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, const int *row_ptr, const int *cand,
                   int *best_idx, int *best_dist, int *second_dist, int *best_pos, int *second_pos) {
    if (!best_idx || !best_dist || !second_dist || !best_pos || !second_pos) return fail(ORBX_ERR_INVALID, "NULL output");
    return match_csr_impl(m, query, nq, train, nt, row_ptr, cand, best_idx, best_dist, second_dist, best_pos, second_pos, nullptr);
}

int orbx_match_csr_dists(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, const int *row_ptr,
                         const int *cand, int *dist) {
    if (!dist) return fail(ORBX_ERR_INVALID, "NULL output");
    return match_csr_impl(m, query, nq, train, nt, row_ptr, cand, nullptr, nullptr, nullptr, nullptr, nullptr, dist);
}

int orbx_search_for_triangulation(orbx_matcher *m, int nn1, const int *nodes1, const int *ptr1, const int *idx1, int nn2,
                                  const int *nodes2, const int *ptr2, const int *idx2, const uint8_t *desc1, int n1,
                                  const uint8_t *desc2, int n2, const uint8_t *skip1, const uint8_t *skip2, int *match12,
                                  int *nmatches) {
    if (!m || !match12 || n1 < 0 || n2 < 0 || nn1 < 0 || nn2 < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    if (nmatches) *nmatches = 0;
    // merge-walk of the two FeatureVectors (src/ORBmatcher.cc:1218-1222, 1420-1427) — host, O(nn1 + nn2)
    std::vector<int> row_q, row_begin, row_len;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int k = ptr1[a]; k < ptr1[a + 1]; k++) {
                const int i1 = idx1[k];
                if (i1 < 0 || i1 >= n1) return fail(ORBX_ERR_INVALID, "idx1 out of range");
                if (skip1 && skip1[i1]) continue;
                row_q.push_back(i1); row_begin.push_back(ptr2[b]); row_len.push_back(ptr2[b + 1] - ptr2[b]);
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) {
            a = (int)(std::lower_bound(nodes1 + a, nodes1 + nn1, nodes2[b]) - nodes1);
        } else {
            b = (int)(std::lower_bound(nodes2 + b, nodes2 + nn2, nodes1[a]) - nodes2);
        }
    }
    const int nrows = (int)row_q.size();
    if (nrows == 0) return ORBX_OK;
    const int tot2 = ptr2[nn2];
    for (int k = 0; k < tot2; k++)
        if (idx2[k] < 0 || idx2[k] >= n2) return fail(ORBX_ERR_INVALID, "idx2 out of range");
    std::vector<uint8_t> zero;
    if (!skip2) { zero.assign(n2, 0); skip2 = zero.data(); }
    int st;
    if ((st = upload(m, m->b[1], desc1, (size_t)n1 * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], desc2, (size_t)n2 * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[9], row_q.data(), (size_t)nrows * sizeof(int))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[10], row_begin.data(), (size_t)nrows * sizeof(int))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[11], row_len.data(), (size_t)nrows * sizeof(int))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[12], idx2, (size_t)tot2 * sizeof(int))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[13], skip2, (size_t)n2)) != ORBX_OK) return st;
    M_CUDA(m->b[14].ensure((size_t)n1 * sizeof(int)));
    M_CUDA(cudaMemsetAsync(m->b[14].p, 0xFF, (size_t)n1 * sizeof(int), m->stream));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_match_groups<<<(nrows + 127) / 128, 128, 0, m->stream>>>((const uint4 *)m->b[1].p, (const uint4 *)m->b[2].p, nrows,
                                                               m->b[9].as<int>(), m->b[10].as<int>(), m->b[11].as<int>(),
                                                               m->b[12].as<int>(), m->b[13].as<uint8_t>(), 50 /* TH_LOW */,
                                                               m->b[14].as<int>());
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(match12, m->b[14].p, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    int nm = 0;
    for (int i = 0; i < n1; i++) nm += match12[i] >= 0;
    if (nmatches) *nmatches = nm;
    return ORBX_OK;
}

// ---- SearchForTriangulation, the whole function incl. bCoarse = false (src/ORBmatcher.cc:1131-1456); kernels in orbx_sft_kernels.cuh
int orbx_search_for_triangulation_full(orbx_matcher *m, int ncam, const int *cam_off1, const orbx_kp *kps1, const uint8_t *desc1,
                                       const uint8_t *has_mp1, const float *u_right1, int nn1, const int *nodes1, const int *ptr1,
                                       const int *idx1, const int *cam_off2, const orbx_kp *kps2, const uint8_t *desc2,
                                       const uint8_t *has_mp2, const float *u_right2, int nn2, const int *nodes2, const int *ptr2,
                                       const int *idx2, const float *ep, const float *sf2, const float *sigma2_1, const float *sigma2_2,
                                       int nlevels, int only_stereo, int coarse, int check_orientation, const orbx_epipolar_pair *pairs,
                                       int *match12, int *nmatches) {
    if (!m || !match12 || !nmatches || !(ncam == 1 || ncam == 2 || ncam == 4) || !cam_off1 || !cam_off2 || nn1 < 0 || nn2 < 0 ||
        nlevels < 1 || nlevels > 16 || !sf2 || !sigma2_1 || !sigma2_2 || !ep || (!coarse && !pairs) || (nn1 && (!nodes1 || !ptr1 || !idx1)) ||
        (nn2 && (!nodes2 || !ptr2 || !idx2)))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    const int n1 = cam_off1[ncam], n2 = cam_off2[ncam];
    if (n1 < 0 || n2 < 0 || (n1 && (!kps1 || !desc1)) || (n2 && (!kps2 || !desc2))) return fail(ORBX_ERR_INVALID, "bad arguments");
    for (int c = 0; c < ncam; c++)
        if (cam_off1[c] > cam_off1[c + 1] || cam_off2[c] > cam_off2[c + 1]) return fail(ORBX_ERR_INVALID, "camera offsets must ascend");
    for (int i = 0; i < n1; i++) if (kps1[i].octave < 0 || kps1[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "kps1: octave out of range");
    for (int i = 0; i < n2; i++) if (kps2[i].octave < 0 || kps2[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "kps2: octave out of range");
    if (!coarse)
        for (int s = (ncam == 1 ? 0 : 1); s < (ncam == 1 ? 1 : ncam == 2 ? 5 : kSftStates); s++)
            if (pairs[s].model != ORBX_CAM_PINHOLE && pairs[s].model != ORBX_CAM_KB8) return fail(ORBX_ERR_INVALID, "pairs: unknown camera model");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    *nmatches = 0;
    // bStereo / skip flags (:1225-1236, 1258-1269)
    std::vector<uint8_t> st1(std::max(n1, 1), 0), st2(std::max(n2, 1), 0), skip2(std::max(n2, 1), 0);
    if (ncam == 1) {
        if (u_right1) for (int i = 0; i < n1; i++) st1[i] = u_right1[i] >= 0;
        if (u_right2) for (int i = 0; i < n2; i++) st2[i] = u_right2[i] >= 0;
    }
    for (int i = 0; i < n2; i++) skip2[i] = (has_mp2 && has_mp2[i]) || (only_stereo && !st2[i]);
    // merge-walk of the two FeatureVectors (:1218-1222, 1420-1427) — host, O(nn1 + nn2); rows in the reference's visiting order
    std::vector<int> row_q, row_begin, row_len;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int k = ptr1[a]; k < ptr1[a + 1]; k++) {
                const int i1 = idx1[k];
                if (i1 < 0 || i1 >= n1) return fail(ORBX_ERR_INVALID, "idx1 out of range");
                if ((has_mp1 && has_mp1[i1]) || (only_stereo && !st1[i1])) continue;
                row_q.push_back(i1); row_begin.push_back(ptr2[b]); row_len.push_back(ptr2[b + 1] - ptr2[b]);
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) {
            a = (int)(std::lower_bound(nodes1 + a, nodes1 + nn1, nodes2[b]) - nodes1);
        } else {
            b = (int)(std::lower_bound(nodes2 + b, nodes2 + nn2, nodes1[a]) - nodes2);
        }
    }
    const int nrows = (int)row_q.size();
    if (nrows == 0) return ORBX_OK;
    const int tot2 = ptr2[nn2];
    for (int k = 0; k < tot2; k++)
        if (idx2[k] < 0 || idx2[k] >= n2) return fail(ORBX_ERR_INVALID, "idx2 out of range");
    std::vector<orbx_epipolar_pair> pz(kSftStates);
    memset(pz.data(), 0, pz.size() * sizeof(orbx_epipolar_pair));
    Packer pk;
    const size_t o_d1 = pk.add(desc1, (size_t)n1 * 32), o_d2 = pk.add(desc2, (size_t)n2 * 32);
    const size_t o_k1 = pk.add(kps1, (size_t)n1 * sizeof(orbx_kp)), o_k2 = pk.add(kps2, (size_t)n2 * sizeof(orbx_kp));
    const size_t o_st1 = pk.add(st1.data(), (size_t)std::max(n1, 1)), o_st2 = pk.add(st2.data(), (size_t)std::max(n2, 1));
    const size_t o_rq = pk.add(row_q.data(), (size_t)nrows * sizeof(int)), o_rb = pk.add(row_begin.data(), (size_t)nrows * sizeof(int));
    const size_t o_rl = pk.add(row_len.data(), (size_t)nrows * sizeof(int)), o_i2 = pk.add(idx2, (size_t)tot2 * sizeof(int));
    const size_t o_sk2 = pk.add(skip2.data(), (size_t)std::max(n2, 1));
    const size_t o_pr = pk.add(pairs ? pairs : pz.data(), kSftStates * sizeof(orbx_epipolar_pair));
    uint8_t *base = nullptr;
    int st;
    if ((st = pk.commit(m, &base)) != ORBX_OK) return st;
    M_CUDA(m->b[14].ensure((size_t)n1 * sizeof(int)));
    M_CUDA(m->b[15].ensure((size_t)nrows * sizeof(int)));
    M_CUDA(m->b[16].ensure((size_t)(nrows + 1) * sizeof(int)));
    M_CUDA(cudaMemsetAsync(m->b[14].p, 0xFF, (size_t)n1 * sizeof(int), m->stream));
    M_CUDA(cudaMemsetAsync(m->b[16].p, 0, (size_t)(nrows + 1) * sizeof(int), m->stream));
    SftArgs A;
    memset(&A, 0, sizeof(A));
    A.d1 = (const uint4 *)(base + o_d1); A.d2 = (const uint4 *)(base + o_d2); A.k1 = (const orbx_kp *)(base + o_k1); A.k2 = (const orbx_kp *)(base + o_k2);
    A.nrows = nrows; A.row_q = (const int *)(base + o_rq); A.row_begin = (const int *)(base + o_rb); A.row_len = (const int *)(base + o_rl);
    A.idx2 = (const int *)(base + o_i2);
    A.skip2 = base + o_sk2; A.st1 = base + o_st1; A.st2 = base + o_st2; A.ncam = ncam;
    for (int c = 0; c <= 4; c++) { A.off1[c] = cam_off1[std::min(c, ncam)]; A.off2[c] = cam_off2[std::min(c, ncam)]; }
    A.ep[0] = ep[0]; A.ep[1] = ep[1];
    for (int l = 0; l < 16; l++) { const int k = std::min(l, nlevels - 1); A.sf2[l] = sf2[k]; A.sigma1[l] = sigma2_1[k]; A.sigma2[l] = sigma2_2[k]; }
    A.coarse = coarse != 0; A.pairs = (const orbx_epipolar_pair *)(base + o_pr);
    A.match12 = m->b[14].as<int>(); A.exit_state = m->b[15].as<int>(); A.unresolved = m->b[16].as<int>();
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_sft_rows<<<(nrows + 127) / 128, 128, 0, m->stream>>>(A);
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    if (ncam >= 2) {     // the carried camera pair only matters when pairs without a branch exist
        int any = 0;
        M_CUDA(cudaMemcpyAsync(&any, m->b[16].as<int>() + nrows, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaStreamSynchronize(m->stream));
        if (any) {
            k_sft_fix<<<1, 32, 0, m->stream>>>(A);
            m->last_launches++;
            M_CUDA(cudaGetLastError());
        }
    }
    if (check_orientation) {
        k_sft_rot<<<1, 256, 0, m->stream>>>(A.k1, A.k2, n1, A.match12);
        m->last_launches++;
        M_CUDA(cudaGetLastError());
    }
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(match12, m->b[14].p, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    int nm = 0;
    for (int i = 0; i < n1; i++) nm += match12[i] >= 0;
    *nmatches = nm;
    return ORBX_OK;
}

// ---- SearchByBoW (src/ORBmatcher.cc:349-666, 1006-1129)
static int search_by_bow_impl(orbx_matcher *m, int variant, int nn1, const int *nodes1, const int *ptr1, const int *idx1, int nn2,
                              const int *nodes2, const int *ptr2, const int *idx2, const uint8_t *desc1, int n1, const uint8_t *desc2,
                              int n2, const uint8_t *skip1, const uint8_t *skip2, const float *angle1, const float *angle2, int ncam2,
                              const int *cam_off2, float nnratio, int check_orientation, int *match, int *nmatches) {
    const int nout = variant == 0 ? n2 : n1;
    if (!m || !match || !nmatches || n1 < 0 || n2 < 0 || nn1 < 0 || nn2 < 0 || (nn1 && (!nodes1 || !ptr1 || !idx1)) ||
        (nn2 && (!nodes2 || !ptr2 || !idx2)) || (n1 && !desc1) || (n2 && !desc2) || (check_orientation && n1 && n2 && (!angle1 || !angle2)) ||
        (variant == 0 && !(ncam2 == 1 || ((ncam2 == 2 || ncam2 == 4) && cam_off2))))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i = 0; i < nout; i++) match[i] = -1;
    *nmatches = 0;
    if (n1 == 0 || n2 == 0 || nn1 == 0 || nn2 == 0) return ORBX_OK;
    if (n1 >= 65535 || n2 >= 65535) return fail(ORBX_ERR_INVALID, "more than 65534 features per frame are not supported");
    const int tot1 = ptr1[nn1], tot2 = ptr2[nn2];
    for (int k = 0; k < tot1; k++)
        if (idx1[k] < 0 || idx1[k] >= n1) return fail(ORBX_ERR_INVALID, "idx1 out of range");
    for (int k = 0; k < tot2; k++)
        if (idx2[k] < 0 || idx2[k] >= n2) return fail(ORBX_ERR_INVALID, "idx2 out of range");
    // merge-walk of the two FeatureVectors (:363-369, 626-634): the shared nodes
    std::vector<int> pa, pb;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) { pa.push_back(a); pb.push_back(b); a++; b++; }
        else if (nodes1[a] < nodes2[b]) a = (int)(std::lower_bound(nodes1 + a, nodes1 + nn1, nodes2[b]) - nodes1);
        else b = (int)(std::lower_bound(nodes2 + b, nodes2 + nn2, nodes1[a]) - nodes2);
    }
    const int npairs = (int)pa.size();
    if (npairs == 0) return ORBX_OK;
    std::vector<uint8_t> s1(n1, 0), s2(n2, 0);
    if (skip1) memcpy(s1.data(), skip1, n1);
    if (skip2) memcpy(s2.data(), skip2, n2);
    std::vector<float> zero;
    if (!check_orientation) { zero.assign(std::max(n1, n2), 0.0f); angle1 = angle2 = zero.data(); }
    Packer pk;
    const size_t o_d1 = pk.add(desc1, (size_t)n1 * 32), o_d2 = pk.add(desc2, (size_t)n2 * 32);
    const size_t o_pa = pk.add(pa.data(), (size_t)npairs * 4), o_pb = pk.add(pb.data(), (size_t)npairs * 4);
    const size_t o_p1 = pk.add(ptr1, (size_t)(nn1 + 1) * 4), o_i1 = pk.add(idx1, (size_t)tot1 * 4);
    const size_t o_p2 = pk.add(ptr2, (size_t)(nn2 + 1) * 4), o_i2 = pk.add(idx2, (size_t)tot2 * 4);
    const size_t o_s1 = pk.add(s1.data(), (size_t)n1), o_s2 = pk.add(s2.data(), (size_t)n2);
    const size_t o_a1 = pk.add(angle1, (size_t)n1 * 4), o_a2 = pk.add(angle2, (size_t)n2 * 4);
    uint8_t *base = nullptr;
    int st;
    if ((st = pk.commit(m, &base)) != ORBX_OK) return st;
    // work: match [nout] | events [4 n1] (both -1) | claimed2 [n2] | nmatches [1] (both 0)
    const size_t o_ev = (size_t)nout, o_cl = o_ev + 4 * (size_t)n1, o_nm = o_cl + n2;
    M_CUDA(m->b[13].ensure((o_nm + 1) * 4));
    int *w = m->b[13].as<int>();
    M_CUDA(cudaMemsetAsync(w, 0xFF, o_cl * 4, m->stream));
    M_CUDA(cudaMemsetAsync(w + o_cl, 0, ((size_t)n2 + 1) * 4, m->stream));
    BowArgs A;
    A.variant = variant; A.npairs = npairs;
    A.pair_a = (const int *)(base + o_pa); A.pair_b = (const int *)(base + o_pb);
    A.ptr1 = (const int *)(base + o_p1); A.idx1 = (const int *)(base + o_i1); A.ptr2 = (const int *)(base + o_p2); A.idx2 = (const int *)(base + o_i2);
    A.desc1 = (const uint4 *)(base + o_d1); A.desc2 = (const uint4 *)(base + o_d2);
    A.skip1 = base + o_s1; A.skip2 = base + o_s2;
    A.angle1 = (const float *)(base + o_a1); A.angle2 = (const float *)(base + o_a2);
    A.ncam2 = variant == 0 ? ncam2 : 1;
    for (int c = 0; c < 5; c++) A.cam_off2[c] = (variant == 0 && ncam2 > 1 && c <= ncam2) ? cam_off2[c] : n2;
    A.nnratio = nnratio; A.check_ori = check_orientation ? 1 : 0;
    A.match = w; A.events = w + o_ev; A.claimed2 = w + o_cl; A.nmatches = w + o_nm;
    A.n1 = n1; A.n2 = n2;
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_bow_match<<<(npairs + 3) / 4, 128, 0, m->stream>>>(A);
    m->last_launches = 1;
    if (check_orientation) {
        k_rot_filter<<<1, 1024, 0, m->stream>>>(A.events, 4 * n1, A.match, A.nmatches);
        m->last_launches++;
    }
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(match, w, (size_t)nout * 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(nmatches, w + o_nm, 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_search_by_bow_frame(orbx_matcher *m, int nn_kf, const int *nodes_kf, const int *ptr_kf, const int *idx_kf, int nn_f,
                             const int *nodes_f, const int *ptr_f, const int *idx_f, const uint8_t *desc_kf, int n_kf,
                             const uint8_t *desc_f, int n_f, const uint8_t *skip_kf, const float *angle_kf, const float *angle_f,
                             int ncam_f, const int *cam_off_f, float nnratio, int check_orientation, int *match_f, int *nmatches) {
    return search_by_bow_impl(m, 0, nn_kf, nodes_kf, ptr_kf, idx_kf, nn_f, nodes_f, ptr_f, idx_f, desc_kf, n_kf, desc_f, n_f, skip_kf, nullptr,
                              angle_kf, angle_f, ncam_f, cam_off_f, nnratio, check_orientation, match_f, nmatches);
}

int orbx_search_by_bow_keyframes(orbx_matcher *m, int nn1, const int *nodes1, const int *ptr1, const int *idx1, int nn2, const int *nodes2,
                                 const int *ptr2, const int *idx2, const uint8_t *desc1, int n1, const uint8_t *desc2, int n2,
                                 const uint8_t *skip1, const uint8_t *skip2, const float *angle1, const float *angle2, float nnratio,
                                 int check_orientation, int *match12, int *nmatches) {
    return search_by_bow_impl(m, 1, nn1, nodes1, ptr1, idx1, nn2, nodes2, ptr2, idx2, desc1, n1, desc2, n2, skip1, skip2, angle1, angle2, 1,
                              nullptr, nnratio, check_orientation, match12, nmatches);
}

// builds the device grid of one camera into buffers b[base .. base+3]; returns device pointers
static int grid_build(orbx_matcher *m, int base, const orbx_kp *d_kps, int n, float min_x, float min_y, float inv_w, float inv_h,
                      int **cell_start, int **cell_items) {
    const int nc = kGridCols * kGridRows;
    M_CUDA(m->b[base].ensure((size_t)nc * sizeof(int)));            // counts
    M_CUDA(m->b[base + 1].ensure((size_t)(nc + 1) * sizeof(int)));  // starts
    M_CUDA(m->b[base + 2].ensure((size_t)nc * sizeof(int)));        // cursors
    M_CUDA(m->b[base + 3].ensure((size_t)std::max(n, 1) * sizeof(int)));
    M_CUDA(cudaMemsetAsync(m->b[base].p, 0, (size_t)nc * sizeof(int), m->stream));
    if (n > 0) k_grid_count<<<(n + 255) / 256, 256, 0, m->stream>>>(d_kps, n, min_x, min_y, inv_w, inv_h, m->b[base].as<int>());
    k_grid_scan<<<1, 1024, 0, m->stream>>>(m->b[base].as<int>(), m->b[base + 1].as<int>(), m->b[base + 2].as<int>());
    if (n > 0) {
        k_grid_fill<<<(n + 255) / 256, 256, 0, m->stream>>>(d_kps, n, min_x, min_y, inv_w, inv_h, m->b[base + 1].as<int>(),
                                                            m->b[base + 2].as<int>(), m->b[base + 3].as<int>());
        k_grid_sort_cells<<<(nc + 255) / 256, 256, 0, m->stream>>>(m->b[base + 1].as<int>(), m->b[base + 3].as<int>());
    }
    m->last_launches += n > 0 ? 4 : 1;
    M_CUDA(cudaGetLastError());
    *cell_start = m->b[base + 1].as<int>();
    *cell_items = m->b[base + 3].as<int>();
    return ORBX_OK;
}

int orbx_grid_candidates(orbx_matcher *m, const orbx_kp *kps, int n, float min_x, float max_x, float min_y, float max_y,
                         const float *qx, const float *qy, const float *qr, const int *min_level, const int *max_level, int nq,
                         int *row_ptr, int *cand, int cand_cap, int *n_cand) {
    if (!m || n < 0 || nq < 0 || !row_ptr || !n_cand || !(max_x > min_x) || !(max_y > min_y)) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    const float inv_w = (float)kGridCols / (max_x - min_x), inv_h = (float)kGridRows / (max_y - min_y);   // Frame.cc:1914-1917
    int st;
    if ((st = upload(m, m->b[1], kps, (size_t)n * sizeof(orbx_kp))) != ORBX_OK) return st;
    int *cell_start, *cell_items;
    if ((st = grid_build(m, 15, m->b[1].as<orbx_kp>(), n, min_x, min_y, inv_w, inv_h, &cell_start, &cell_items)) != ORBX_OK) return st;
    row_ptr[0] = 0; *n_cand = 0;
    if (nq == 0) { M_CUDA(cudaStreamSynchronize(m->stream)); return ORBX_OK; }
    const size_t fb = (size_t)nq * sizeof(float), ib = (size_t)nq * sizeof(int);
    M_CUDA(m->b[5].ensure(3 * fb + 2 * ib));
    float *dqx = m->b[5].as<float>(), *dqy = dqx + nq, *dqr = dqy + nq;
    int *dmin = (int *)(dqr + nq), *dmax = dmin + nq;
    M_CUDA(cudaMemcpyAsync(dqx, qx, fb, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(dqy, qy, fb, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(dqr, qr, fb, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(dmin, min_level, ib, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(dmax, max_level, ib, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(m->b[6].ensure(ib));
    M_CUDA(m->b[7].ensure(ib + sizeof(int)));
    M_CUDA(m->b[8].ensure((size_t)std::max(cand_cap, 1) * sizeof(int)));
    const orbx_kp *dk = m->b[1].as<orbx_kp>();
    k_grid_query_count<<<(nq + 127) / 128, 128, 0, m->stream>>>(dk, cell_start, cell_items, min_x, min_y, inv_w, inv_h, dqx, dqy, dqr,
                                                                dmin, dmax, nq, m->b[6].as<int>());
    k_excl_scan<<<1, 1024, 0, m->stream>>>(m->b[6].as<int>(), nq, m->b[7].as<int>());
    k_grid_query_fill<<<(nq + 127) / 128, 128, 0, m->stream>>>(dk, cell_start, cell_items, min_x, min_y, inv_w, inv_h, dqx, dqy, dqr,
                                                               dmin, dmax, nq, m->b[7].as<int>(), m->b[8].as<int>(), cand_cap);
    m->last_launches += 3;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaMemcpyAsync(row_ptr, m->b[7].p, ib + sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    *n_cand = row_ptr[nq];
    if (*n_cand > cand_cap) return fail(ORBX_ERR_CAPACITY, "candidate buffer too small");
    if (*n_cand > 0 && cand) {
        M_CUDA(cudaMemcpyAsync(cand, m->b[8].p, (size_t)*n_cand * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        M_CUDA(cudaStreamSynchronize(m->stream));
    }
    return ORBX_OK;
}

// ---- window search without claims (Fuse, SearchBySim3)
// device-side: grid of `n` uploaded keypoints in b[base .. base + 3], results in d_idx / d_dist (device)
static int window_best_device(orbx_matcher *m, int base, const orbx_kp *d_kps, const uint4 *d_desc, int n, const float *gb, const uint8_t *d_skip,
                              const float *d_uright, const float *d_inv_sigma2, int nq, const uint8_t *d_valid, const float *d_qx,
                              const float *d_qy, const float *d_qur, const float *d_radius, const int *d_min, const int *d_max,
                              const uint4 *d_qdesc, int *d_idx, int *d_dist) {
    WindowArgs A;
    A.kps = d_kps; A.desc = d_desc;
    A.min_x = gb[0]; A.min_y = gb[2];
    A.inv_w = (float)kGridCols / (gb[1] - gb[0]); A.inv_h = (float)kGridRows / (gb[3] - gb[2]);
    int *cs, *ci, st;
    if ((st = grid_build(m, base, d_kps, n, A.min_x, A.min_y, A.inv_w, A.inv_h, &cs, &ci)) != ORBX_OK) return st;
    A.cell_start = cs; A.cell_items = ci;
    A.skip = d_skip; A.u_right = d_uright; A.inv_sigma2 = d_inv_sigma2;
    A.nq = nq; A.valid = d_valid; A.qx = d_qx; A.qy = d_qy; A.q_ur = d_qur; A.radius = d_radius; A.min_level = d_min; A.max_level = d_max;
    A.q_desc = d_qdesc; A.best_idx = d_idx; A.best_dist = d_dist;
    k_window_best<<<(nq + 127) / 128, 128, 0, m->stream>>>(A);
    m->last_launches++;
    M_CUDA(cudaGetLastError());
    return ORBX_OK;
}

int orbx_window_best(orbx_matcher *m, const orbx_kp *kps, const uint8_t *desc, int n, const float *grid_bounds, const uint8_t *skip,
                     const float *u_right, const float *inv_level_sigma2, int nlevels, int nq, const uint8_t *valid, const float *qx,
                     const float *qy, const float *q_ur, const float *radius, const int *min_level, const int *max_level,
                     const uint8_t *q_desc, int *best_idx, int *best_dist) {
    if (!m || n < 0 || nq < 0 || !grid_bounds || !(grid_bounds[1] > grid_bounds[0]) || !(grid_bounds[3] > grid_bounds[2]) ||
        (n && (!kps || !desc)) || (nq && (!valid || !qx || !qy || !radius || !min_level || !max_level || !q_desc || !best_idx || !best_dist)) ||
        (u_right && (!q_ur || !inv_level_sigma2)) || (inv_level_sigma2 && (nlevels < 1 || nlevels > 32)))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (nq == 0) return ORBX_OK;
    if (inv_level_sigma2)
        for (int i = 0; i < n; i++)
            if (kps[i].octave < 0 || kps[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "keypoint octave outside the level table");
    int st;
    if ((st = upload(m, m->b[1], kps, (size_t)n * sizeof(orbx_kp))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], desc, (size_t)n * 32)) != ORBX_OK) return st;
    const size_t fb = (size_t)nq * 4;
    M_CUDA(m->b[5].ensure(16 * 16 + 6 * fb + (size_t)nq * 33 + (size_t)n * 5 + 32 * 4));
    uint8_t *base = m->b[5].as<uint8_t>();
    size_t off = 0;
    auto put = [&](const void *src, size_t nb) -> void * {
        if (!src) return nullptr;
        void *d = base + off;
        cudaMemcpyAsync(d, src, nb, cudaMemcpyHostToDevice, m->stream);
        off += (nb + 15) & ~(size_t)15;
        return d;
    };
    const float *d_qx = (const float *)put(qx, fb), *d_qy = (const float *)put(qy, fb), *d_qur = (const float *)put(q_ur, fb);
    const float *d_r = (const float *)put(radius, fb);
    const int *d_mn = (const int *)put(min_level, fb), *d_mx = (const int *)put(max_level, fb);
    const uint4 *d_qd = (const uint4 *)put(q_desc, (size_t)nq * 32);
    const uint8_t *d_valid = (const uint8_t *)put(valid, nq), *d_skip = (const uint8_t *)put(skip, n);
    const float *d_ur = (const float *)put(u_right, (size_t)n * 4), *d_is = (const float *)put(inv_level_sigma2, (size_t)nlevels * 4);
    M_CUDA(cudaGetLastError());
    M_CUDA(m->b[6].ensure(2 * fb));
    int *d_idx = m->b[6].as<int>(), *d_dist = d_idx + nq;
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    st = window_best_device(m, 15, m->b[1].as<orbx_kp>(), (const uint4 *)m->b[2].p, n, grid_bounds, d_skip, d_ur, d_is, nq, d_valid, d_qx, d_qy,
                            d_qur, d_r, d_mn, d_mx, d_qd, d_idx, d_dist);
    if (st != ORBX_OK) return st;
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(best_idx, d_idx, fb, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(best_dist, d_dist, fb, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_search_by_sim3(orbx_matcher *m, const orbx_kp *kps1, const uint8_t *desc1, int n1, const float *grid_bounds1,
                        const orbx_kp *kps2, const uint8_t *desc2, int n2, const float *grid_bounds2, const uint8_t *valid1,
                        const float *u1, const float *v1, const float *radius1, const int *level1, const uint8_t *mp_desc1,
                        const uint8_t *valid2, const float *u2, const float *v2, const float *radius2, const int *level2,
                        const uint8_t *mp_desc2, int *match12, int *nfound) {
    if (!m || n1 < 0 || n2 < 0 || !match12 || !nfound || (n1 && (!kps1 || !desc1 || !valid1 || !u1 || !v1 || !radius1 || !level1 || !mp_desc1)) ||
        (n2 && (!kps2 || !desc2 || !valid2 || !u2 || !v2 || !radius2 || !level2 || !mp_desc2)) || !grid_bounds1 || !grid_bounds2)
        return fail(ORBX_ERR_INVALID, "bad arguments");
    for (int i = 0; i < n1; i++) match12[i] = -1;
    *nfound = 0;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    std::vector<int> mn1(n1), mn2(n2), b1(n1), d1(n1), b2(n2), d2(n2);
    for (int i = 0; i < n1; i++) mn1[i] = level1[i] - 1;      // kp.octave < nPredictedLevel - 1 || kp.octave > nPredictedLevel (:1851, 1925)
    for (int i = 0; i < n2; i++) mn2[i] = level2[i] - 1;
    int st;
    // KF1's map points searched in KF2 (:1805-1866), then KF2's in KF1 (:1869-1940)
    if ((st = orbx_window_best(m, kps2, desc2, n2, grid_bounds2, nullptr, nullptr, nullptr, 0, n1, valid1, u1, v1, nullptr, radius1, mn1.data(),
                               level1, mp_desc1, b1.data(), d1.data())) != ORBX_OK) return st;
    const int l1 = m->last_launches;
    if ((st = orbx_window_best(m, kps1, desc1, n1, grid_bounds1, nullptr, nullptr, nullptr, 0, n2, valid2, u2, v2, nullptr, radius2, mn2.data(),
                               level2, mp_desc2, b2.data(), d2.data())) != ORBX_OK) return st;
    m->last_launches += l1;
    int nf = 0;
    for (int i1 = 0; i1 < n1; i1++) {   // agreement (:1968-1980), TH_HIGH = 100 (:1863, 1937)
        const int idx2 = d1[i1] <= 100 ? b1[i1] : -1;
        if (idx2 < 0) continue;
        const int idx1 = d2[idx2] <= 100 ? b2[idx2] : -1;
        if (idx1 == i1) { match12[i1] = idx2; nf++; }
    }
    *nfound = nf;
    return ORBX_OK;
}

// frame side of the projection matchers: keypoints, descriptors, one 64x48 grid per camera, occupancy, stereo mirror tables
static int sbp_setup_frame(orbx_matcher *m, int ncam, const int *cam_off, const orbx_kp *kps, const uint8_t *desc, const float *grid_bounds,
                           const float *scale_factors, int nlevels, const uint8_t *occupied, const int *l2r, const int *r2l, SbpFrame &F) {
    const int N = cam_off[ncam];
    int st;
    if ((st = upload(m, m->b[1], kps, (size_t)N * sizeof(orbx_kp))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], desc, (size_t)N * 32)) != ORBX_OK) return st;
    memset(&F, 0, sizeof(F));
    F.ncam = ncam; F.n_total = N;
    F.kps = m->b[1].as<orbx_kp>(); F.desc = (const uint4 *)m->b[2].p;
    for (int c = 0; c <= ncam; c++) F.cam_off[c] = cam_off[c];
    for (int c = 0; c < ncam; c++) {
        const float *gb = grid_bounds + 4 * c;
        if (!(gb[1] > gb[0]) || !(gb[3] > gb[2])) return fail(ORBX_ERR_INVALID, "bad grid bounds");
        F.min_x[c] = gb[0]; F.min_y[c] = gb[2];
        F.inv_w[c] = (float)kGridCols / (gb[1] - gb[0]); F.inv_h[c] = (float)kGridRows / (gb[3] - gb[2]);
    }
    // grids: 4 cameras x {count,start,cursor,items} packed into b[15..18] with per-camera offsets
    const int nc = kGridCols * kGridRows;
    M_CUDA(m->b[15].ensure((size_t)4 * nc * sizeof(int)));
    M_CUDA(m->b[16].ensure((size_t)4 * (nc + 1) * sizeof(int)));
    M_CUDA(m->b[17].ensure((size_t)4 * nc * sizeof(int)));
    M_CUDA(m->b[18].ensure((size_t)N * sizeof(int)));
    M_CUDA(cudaMemsetAsync(m->b[15].p, 0, (size_t)4 * nc * sizeof(int), m->stream));
    for (int c = 0; c < ncam; c++) {
        const int n = cam_off[c + 1] - cam_off[c];
        const orbx_kp *dk = F.kps + cam_off[c];
        int *cnt = m->b[15].as<int>() + c * nc, *start = m->b[16].as<int>() + c * (nc + 1), *cur = m->b[17].as<int>() + c * nc;
        int *items = m->b[18].as<int>() + cam_off[c];
        if (n > 0) k_grid_count<<<(n + 255) / 256, 256, 0, m->stream>>>(dk, n, F.min_x[c], F.min_y[c], F.inv_w[c], F.inv_h[c], cnt);
        k_grid_scan<<<1, 1024, 0, m->stream>>>(cnt, start, cur);
        if (n > 0) {
            k_grid_fill<<<(n + 255) / 256, 256, 0, m->stream>>>(dk, n, F.min_x[c], F.min_y[c], F.inv_w[c], F.inv_h[c], start, cur, items);
            k_grid_sort_cells<<<(nc + 255) / 256, 256, 0, m->stream>>>(start, items);
        }
        m->last_launches += n > 0 ? 4 : 1;
        F.cell_start[c] = start; F.cell_items[c] = items;
    }
    if ((st = upload(m, m->b[19], scale_factors, (size_t)nlevels * sizeof(float))) != ORBX_OK) return st;
    F.scale_factors = m->b[19].as<float>();
    // occupied / l2r / r2l
    std::vector<uint8_t> occ(N, 0);
    if (occupied) memcpy(occ.data(), occupied, N);
    if ((st = upload(m, m->b[20], occ.data(), (size_t)N)) != ORBX_OK) return st;
    F.occupied0 = m->b[20].as<uint8_t>();
    const int nleft = cam_off[1] - cam_off[0], nright = ncam >= 2 ? cam_off[2] - cam_off[1] : 0;
    std::vector<int> lr(std::max(nleft, 1), -1), rl(std::max(nright, 1), -1);
    if (l2r && ncam >= 2) memcpy(lr.data(), l2r, (size_t)nleft * sizeof(int));
    if (r2l && ncam >= 2) memcpy(rl.data(), r2l, (size_t)nright * sizeof(int));
    M_CUDA(m->b[21].ensure((lr.size() + rl.size()) * sizeof(int)));
    M_CUDA(cudaMemcpyAsync(m->b[21].p, lr.data(), lr.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(m->b[21].as<int>() + lr.size(), rl.data(), rl.size() * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    F.l2r = m->b[21].as<int>(); F.r2l = m->b[21].as<int>() + lr.size();
    return ORBX_OK;
}

int orbx_search_by_projection(orbx_matcher *m, int ncam, const int *cam_off, const orbx_kp *kps, const uint8_t *desc,
                              const float *grid_bounds, const float *scale_factors, int nlevels, const uint8_t *occupied,
                              const int *l2r, const int *r2l, const float *u_right, int nmp, const uint8_t *in_view, const float *proj_x,
                              const float *proj_y, const int *level, const float *view_cos, const float *proj_xr, const float *depth, const uint8_t *bad,
                              const uint8_t *has_obs, const uint8_t *mp_desc, float th, int far_points, float th_far_points,
                              float nnratio, int *assigned, int *nmatches) {
    if (!m || !cam_off || !kps || !desc || !grid_bounds || !scale_factors || !assigned || !nmatches || nmp < 0 ||
        !(ncam == 1 || ncam == 2 || ncam == 4) || nlevels < 1 || nlevels > 32)
        return fail(ORBX_ERR_INVALID, "bad arguments");
    if (nmp > 0 && (!in_view || !proj_x || !proj_y || !level || !view_cos || !mp_desc)) return fail(ORBX_ERR_INVALID, "NULL map-point array");
    if ((u_right != nullptr) != (proj_xr != nullptr)) return fail(ORBX_ERR_INVALID, "u_right and proj_xr go together");
    for (int i = 0; i < nmp * 4; i++)      // sbp_query indexes scale_factors[level]: validate what it will read
        if (in_view[i] && (i & 3) < ncam && !((i & 3) >= 1 && level[i] == -1) && (level[i] < 0 || level[i] >= nlevels))
            return fail(ORBX_ERR_INVALID, "predicted level out of range");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    const int N = cam_off[ncam];
    for (int i = 0; i < N; i++) assigned[i] = -1;
    *nmatches = 0;
    if (nmp == 0 || N == 0) return ORBX_OK;
    if (N >= 65535) return fail(ORBX_ERR_INVALID, "more than 65534 keypoints per frame are not supported");
    int st;
    SbpFrame F;
    if ((st = sbp_setup_frame(m, ncam, cam_off, kps, desc, grid_bounds, scale_factors, nlevels, occupied, l2r, r2l, F)) != ORBX_OK) return st;
    // ---- map-point side: one query per (map point, camera) in reference order
    const size_t q = (size_t)nmp * 4;
    const size_t bytes = q * (1 + 4 + 4 + 4 + 4) + (size_t)nmp * (4 + 4 + 1 + 1 + 32) + (size_t)N * 4;
    uint8_t *base = nullptr;
    SbpPoints P;
    P.nmp = nmp;
    size_t off = 0;
    cudaError_t put_err = cudaSuccess;
    auto put = [&](const void *src, size_t nb) -> void * {
        void *d = base + off;
        const cudaError_t e = cudaMemcpyAsync(d, src, nb, cudaMemcpyHostToDevice, m->stream);
        if (e != cudaSuccess && put_err == cudaSuccess) put_err = e;
        off += (nb + 15) & ~(size_t)15;
        return d;
    };
    M_CUDA(m->b[22].ensure(bytes + 16 * 16));
    base = m->b[22].as<uint8_t>();
    P.proj_x = (const float *)put(proj_x, q * 4);
    P.proj_y = (const float *)put(proj_y, q * 4);
    P.level = (const int *)put(level, q * 4);
    P.view_cos = (const float *)put(view_cos, q * 4);
    P.desc = (const uint4 *)put(mp_desc, (size_t)nmp * 32);
    P.in_view = (const uint8_t *)put(in_view, q);
    std::vector<uint8_t> active(nmp), hobs(nmp);
    for (int i = 0; i < nmp; i++) {   // the per-point early `continue`s (src/ORBmatcher.cc:34-43)
        const uint8_t *iv = in_view + 4 * i;
        bool a = iv[0] || iv[1] || iv[2] || iv[3];
        if (far_points && depth && depth[i] > th_far_points) a = false;
        if (bad && bad[i]) a = false;
        active[i] = a; hobs[i] = has_obs ? has_obs[i] : 1;
    }
    P.active = (const uint8_t *)put(active.data(), nmp);
    P.has_obs = (const uint8_t *)put(hobs.data(), nmp);
    P.u_right = P.proj_ur = nullptr;
    if (u_right && ncam == 1) {               // rectified-stereo gate (:81-85)
        P.u_right = (const float *)put(u_right, (size_t)N * 4);
        P.proj_ur = (const float *)put(proj_xr, (size_t)nmp * 4);
    }
    M_CUDA(put_err);
    M_CUDA(cudaGetLastError());
    P.th = th; P.nnratio = nnratio;
    P.mode = 0; P.level_gate = 0; P.check_ori = 0; P.max_dist = 100; P.angle = nullptr;
    // ---- device work buffers
    M_CUDA(m->b[23].ensure(sbp_work_bytes(nmp, N)));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    st = sbp_run(F, P, m->b[23].as<uint8_t>(), m->b[0], m->stream, &m->last_launches);
    if (st != ORBX_OK) return st;
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    SbpResult R;
    sbp_result_ptrs(m->b[23].as<uint8_t>(), nmp, N, &R);
    M_CUDA(cudaMemcpyAsync(assigned, R.assigned, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    int nm_st[2] = {0, 0};
    M_CUDA(cudaMemcpyAsync(nm_st, R.nmatches, 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    *nmatches = nm_st[0];
    if (nm_st[1] != 0) return fail(ORBX_ERR_CUDA, "ordered resolver did not converge (internal error)");
    return ORBX_OK;
}

int orbx_search_by_projection_last(orbx_matcher *m, int ncam, const int *cam_off, const orbx_kp *kps, const uint8_t *desc,
                                   const float *grid_bounds, const float *scale_factors, int nlevels, const uint8_t *occupied,
                                   const float *u_right, int npts, const uint8_t *valid, const float *proj_x, const float *proj_y,
                                   const float *proj_ur, const int *last_octave, const float *last_angle, const uint8_t *has_obs,
                                   const uint8_t *pt_desc, float th, int level_gate, int check_orientation, int max_dist,
                                   int *assigned, int *nmatches) {
    if (!m || !cam_off || !kps || !desc || !grid_bounds || !scale_factors || !assigned || !nmatches || npts < 0 ||
        !(ncam == 1 || ncam == 2 || ncam == 4) || nlevels < 1 || nlevels > 32 || max_dist < 0 || max_dist > 255 || level_gate < 0 || level_gate > 3 || (level_gate == ORBX_GATE_BELOW && ncam != 1) ||
        (npts && (!valid || !proj_x || !proj_y || !last_octave || !last_angle || !pt_desc)) || (u_right && !proj_ur))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    const int N = cam_off[ncam];
    for (int i = 0; i < N; i++) assigned[i] = -1;
    *nmatches = 0;
    m->last_cleared.assign((size_t)std::max(N, 0), 0);
    if (npts == 0 || N == 0) return ORBX_OK;
    if (N >= 65535) return fail(ORBX_ERR_INVALID, "more than 65534 keypoints per frame are not supported");
    for (int i = 0; i < npts; i++)
        if (valid[i] && (last_octave[i] < 0 || last_octave[i] >= nlevels)) return fail(ORBX_ERR_INVALID, "keypoint octave outside the level table");
    int st;
    SbpFrame F;
    if ((st = sbp_setup_frame(m, ncam, cam_off, kps, desc, grid_bounds, scale_factors, nlevels, occupied, nullptr, nullptr, F)) != ORBX_OK) return st;
    const size_t q = (size_t)npts * 4;
    const size_t bytes = q * (1 + 4 + 4 + 4) + (size_t)npts * (4 + 4 + 1 + 1 + 32) + (size_t)N * 4;
    M_CUDA(m->b[22].ensure(bytes + 16 * 16));
    uint8_t *base = m->b[22].as<uint8_t>();
    size_t off = 0;
    auto put = [&](const void *src, size_t nb) -> void * {
        void *d = base + off;
        cudaMemcpyAsync(d, src, nb, cudaMemcpyHostToDevice, m->stream);
        off += (nb + 15) & ~(size_t)15;
        return d;
    };
    // per-(point, camera) arrays of the shared machinery: every camera of a point uses the octave of the last keypoint and the
    // point's validity (the reference tests the image bounds on camera 0 only, src/ORBmatcher.cc:2021-2026)
    std::vector<int> lvl4(q);
    std::vector<uint8_t> view4(q), act(npts), hobs(npts);
    for (int i = 0; i < npts; i++) {
        for (int c = 0; c < 4; c++) { lvl4[4 * (size_t)i + c] = last_octave[i]; view4[4 * (size_t)i + c] = valid[i] ? 1 : 0; }
        act[i] = valid[i] ? 1 : 0; hobs[i] = has_obs ? has_obs[i] : 1;
    }
    SbpPoints P;
    P.nmp = npts;
    P.proj_x = (const float *)put(proj_x, q * 4);
    P.proj_y = (const float *)put(proj_y, q * 4);
    P.level = (const int *)put(lvl4.data(), q * 4);
    P.view_cos = nullptr;
    P.desc = (const uint4 *)put(pt_desc, (size_t)npts * 32);
    P.in_view = (const uint8_t *)put(view4.data(), q);
    P.active = (const uint8_t *)put(act.data(), npts);
    P.has_obs = (const uint8_t *)put(hobs.data(), npts);
    P.angle = (const float *)put(last_angle, (size_t)npts * 4);
    P.u_right = P.proj_ur = nullptr;
    if (u_right && ncam == 1) {
        P.u_right = (const float *)put(u_right, (size_t)N * 4);
        P.proj_ur = (const float *)put(proj_ur, (size_t)npts * 4);
    }
    M_CUDA(cudaGetLastError());
    P.th = th; P.nnratio = 0.0f;
    P.mode = 1; P.level_gate = level_gate; P.check_ori = check_orientation ? 1 : 0; P.max_dist = max_dist;
    M_CUDA(m->b[23].ensure(sbp_work_bytes(npts, N)));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    st = sbp_run(F, P, m->b[23].as<uint8_t>(), m->b[0], m->stream, &m->last_launches);
    if (st != ORBX_OK) return st;
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    SbpResult R;
    sbp_result_ptrs(m->b[23].as<uint8_t>(), npts, N, &R);
    M_CUDA(cudaMemcpyAsync(assigned, R.assigned, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    int nm_st[2] = {0, 0};
    M_CUDA(cudaMemcpyAsync(nm_st, R.nmatches, 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    *nmatches = nm_st[0];
    if (nm_st[1] != 0) return fail(ORBX_ERR_CUDA, "ordered resolver did not converge (internal error)");
    m->last_cleared.assign((size_t)N, 0);
    for (int i = 0; i < N; i++)
        if (assigned[i] == -2) { assigned[i] = -1; m->last_cleared[i] = 1; }
    return ORBX_OK;
}

int orbx_search_by_projection_last_cleared(const orbx_matcher *m, uint8_t *cleared, int n) {
    if (!m || !cleared || n < 0) return fail(ORBX_ERR_INVALID, "orbx_search_by_projection_last_cleared: bad argument");
    if ((size_t)n != m->last_cleared.size()) return fail(ORBX_ERR_INVALID, "orbx_search_by_projection_last_cleared: n differs from the last call's keypoint count");
    if (n) memcpy(cleared, m->last_cleared.data(), (size_t)n);
    return ORBX_OK;
}

int orbx_search_for_initialization(orbx_matcher *m, const orbx_kp *kps1, const uint8_t *desc1, int n1, const orbx_kp *kps2,
                                   const uint8_t *desc2, int n2, const float *grid_bounds2, float *prev_matched, int window_size,
                                   float nnratio, int check_orientation, int *matches12, int *nmatches) {
    if (!m || n1 < 0 || n2 < 0 || !matches12 || !nmatches || !grid_bounds2 || (n1 && (!kps1 || !desc1 || !prev_matched)) ||
        (n2 && (!kps2 || !desc2)))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    *nmatches = 0;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    if (n1 >= 65535 || n2 >= 65535) return fail(ORBX_ERR_INVALID, "more than 65534 keypoints per frame are not supported");
    int st;
    SbpFrame F;
    const int cam_off[2] = {0, n2};
    const float sf1[1] = {1.0f};
    if ((st = sbp_setup_frame(m, 1, cam_off, kps2, desc2, grid_bounds2, sf1, 1, nullptr, nullptr, nullptr, F)) != ORBX_OK) return st;
    const size_t q = (size_t)n1 * 4;
    M_CUDA(m->b[22].ensure(q * (1 + 4 + 4 + 4) + (size_t)n1 * (4 + 1 + 1 + 32) + 16 * 16));
    uint8_t *base = m->b[22].as<uint8_t>();
    size_t off = 0;
    auto put = [&](const void *src, size_t nb) -> void * {
        void *d = base + off;
        cudaMemcpyAsync(d, src, nb, cudaMemcpyHostToDevice, m->stream);
        off += (nb + 15) & ~(size_t)15;
        return d;
    };
    std::vector<float> px(q, 0.0f), py(q, 0.0f), ang(n1);
    std::vector<int> lvl4(q, 0);
    std::vector<uint8_t> view4(q, 0), act(n1), hobs(n1, 1);
    for (int i = 0; i < n1; i++) {
        px[4 * (size_t)i] = prev_matched[2 * i]; py[4 * (size_t)i] = prev_matched[2 * i + 1];
        lvl4[4 * (size_t)i] = kps1[i].octave;
        act[i] = kps1[i].octave > 0 ? 0 : 1;                       // `if (level1 > 0) continue;` (:912-913)
        view4[4 * (size_t)i] = act[i] && kps1[i].octave == 0;       // (a negative octave finds nothing: kp.octave < minLevel never false... gate [l, l])
        ang[i] = kps1[i].angle;
    }
    SbpPoints P;
    P.nmp = n1;
    P.proj_x = (const float *)put(px.data(), q * 4);
    P.proj_y = (const float *)put(py.data(), q * 4);
    P.level = (const int *)put(lvl4.data(), q * 4);
    P.view_cos = nullptr;
    P.desc = (const uint4 *)put(desc1, (size_t)n1 * 32);
    P.in_view = (const uint8_t *)put(view4.data(), q);
    P.active = (const uint8_t *)put(act.data(), n1);
    P.has_obs = (const uint8_t *)put(hobs.data(), n1);
    P.angle = (const float *)put(ang.data(), (size_t)n1 * 4);
    P.u_right = P.proj_ur = nullptr;
    M_CUDA(cudaGetLastError());
    P.th = (float)window_size; P.nnratio = nnratio;
    P.mode = 2; P.level_gate = 0; P.check_ori = check_orientation ? 1 : 0; P.max_dist = 50;
    M_CUDA(m->b[23].ensure(sbp_work_bytes(n1, n2)));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    st = sbp_run(F, P, m->b[23].as<uint8_t>(), m->b[0], m->stream, &m->last_launches);
    if (st != ORBX_OK) return st;
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    SbpResult R;
    sbp_result_ptrs(m->b[23].as<uint8_t>(), n1, n2, &R);
    M_CUDA(cudaMemcpyAsync(matches12, R.match12, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    int nm_st[2] = {0, 0};
    M_CUDA(cudaMemcpyAsync(nm_st, R.nmatches, 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    *nmatches = nm_st[0];
    if (nm_st[1] != 0) return fail(ORBX_ERR_CUDA, "ordered resolver did not converge (internal error)");
    for (int i = 0; i < n1; i++)                                    // update prev matched (:992-995)
        if (matches12[i] >= 0) { prev_matched[2 * i] = kps2[matches12[i]].x; prev_matched[2 * i + 1] = kps2[matches12[i]].y; }
    return ORBX_OK;
}

// ---- fisheye stereo post-filter
int orbx_kb8_triangulate_matches(orbx_matcher *m, const orbx_kb8_stereo *rig, const float *xy1, const float *xy2, const float *sigma1,
                                 const float *sigma2, int n, float *depth, float *p3d) {
    if (!m || !rig || n < 0 || (n && (!xy1 || !xy2 || !sigma1 || !sigma2 || !depth || !p3d))) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (n == 0) return ORBX_OK;
    int st;
    if ((st = upload(m, m->b[1], xy1, (size_t)n * 2 * sizeof(float))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], xy2, (size_t)n * 2 * sizeof(float))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[3], sigma1, (size_t)n * sizeof(float))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[4], sigma2, (size_t)n * sizeof(float))) != ORBX_OK) return st;
    M_CUDA(m->b[5].ensure((size_t)n * sizeof(float)));
    M_CUDA(m->b[6].ensure((size_t)n * 3 * sizeof(float)));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_kb8_triangulate<<<(n + 127) / 128, 128, 0, m->stream>>>(*rig, m->b[1].as<float>(), m->b[2].as<float>(), m->b[3].as<float>(),
                                                            m->b[4].as<float>(), n, m->b[5].as<float>(), m->b[6].as<float>());
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(depth, m->b[5].p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(p3d, m->b[6].p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

// device core: keypoints / descriptors already on the device (all rows of both cameras; the lapping blocks start at mono_*)
static int fisheye_core(orbx_matcher *m, const orbx_kb8_stereo *rig, const orbx_kp *d_kps_left, const uint8_t *d_desc_left, int n_left,
                        int mono_left, const orbx_kp *d_kps_right, const uint8_t *d_desc_right, int n_right, int mono_right,
                        const float *level_sigma2, int nlevels, double ratio, int *left_to_right, int *right_to_left, float *depth,
                        float *p3d, int *counts) {
    for (int i = 0; i < n_left; i++) { left_to_right[i] = -1; depth[i] = -1.0f; p3d[3 * i] = p3d[3 * i + 1] = p3d[3 * i + 2] = 0.0f; }
    for (int i = 0; i < n_right; i++) right_to_left[i] = -1;
    if (counts) counts[0] = counts[1] = 0;
    const int nq = n_left - mono_left, nt = n_right - mono_right;
    if (nq == 0 || nt < 2) return ORBX_OK;                                            // knnMatch returns fewer than 2 neighbours
    M_CUDA(m->b[3].ensure((size_t)nq * 2 * sizeof(int)));
    M_CUDA(m->b[4].ensure((size_t)nq * 2 * sizeof(int)));
    // outputs: l2r [n_left] | r2l [n_right] | depth [n_left] | p3d [3 n_left] | counters [2], pre-filled from the host copies
    const size_t o_r2l = (size_t)n_left, o_depth = o_r2l + n_right, o_p3d = o_depth + n_left, o_cnt = o_p3d + 3 * (size_t)n_left;
    M_CUDA(m->b[7].ensure((o_cnt + 2) * 4));
    int *d_out = m->b[7].as<int>();
    M_CUDA(cudaMemcpyAsync(d_out, left_to_right, (size_t)n_left * 4, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(d_out + o_r2l, right_to_left, (size_t)n_right * 4, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemcpyAsync(d_out + o_depth, depth, (size_t)n_left * 4, cudaMemcpyHostToDevice, m->stream));
    M_CUDA(cudaMemsetAsync(d_out + o_p3d, 0, (3 * (size_t)n_left + 2) * 4, m->stream));
    int st = orbx_knn2_device(m, d_desc_left + (size_t)mono_left * 32, nq, d_desc_right + (size_t)mono_right * 32, nt, m->b[3].as<int>(),
                              m->b[4].as<int>());
    if (st != ORBX_OK) return st;
    FisheyeArgs A;
    A.kpL = d_kps_left; A.kpR = d_kps_right;
    A.nL = n_left; A.nR = n_right; A.monoL = mono_left; A.monoR = mono_right;
    A.idx = m->b[3].as<int>(); A.dist = m->b[4].as<int>();
    for (int l = 0; l < 16; l++) A.sigma2[l] = l < nlevels ? level_sigma2[l] : 0.0f;
    A.ratio = ratio;
    A.l2r = d_out; A.r2l = d_out + o_r2l; A.depth = (float *)(d_out + o_depth); A.p3d = (float *)(d_out + o_p3d); A.counters = d_out + o_cnt;
    k_fisheye_matches<<<(nq + 127) / 128, 128, 0, m->stream>>>(*rig, A);
    m->last_launches++;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    M_CUDA(cudaMemcpyAsync(left_to_right, d_out, (size_t)n_left * 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(right_to_left, d_out + o_r2l, (size_t)n_right * 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(depth, d_out + o_depth, (size_t)n_left * 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(p3d, d_out + o_p3d, 3 * (size_t)n_left * 4, cudaMemcpyDeviceToHost, m->stream));
    int cnt[2] = {0, 0};
    M_CUDA(cudaMemcpyAsync(cnt, d_out + o_cnt, 2 * 4, cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    if (counts) { counts[0] = cnt[0]; counts[1] = cnt[1]; }
    return ORBX_OK;
}

static bool fisheye_args_ok(orbx_matcher *m, const orbx_kb8_stereo *rig, const void *kl, const void *dl, int n_left, int mono_left,
                            const void *kr, const void *dr, int n_right, int mono_right, const float *level_sigma2, int nlevels,
                            const int *l2r, const int *r2l, const float *depth, const float *p3d) {
    return m && rig && n_left >= 0 && n_right >= 0 && mono_left >= 0 && mono_left <= n_left && mono_right >= 0 && mono_right <= n_right &&
           level_sigma2 && nlevels >= 1 && nlevels <= 16 && (!n_left || (kl && dl && l2r && depth && p3d)) && (!n_right || (kr && dr && r2l));
}

int orbx_stereo_fisheye_matches(orbx_matcher *m, const orbx_kb8_stereo *rig, const orbx_kp *kps_left, const uint8_t *desc_left,
                                int n_left, int mono_left, const orbx_kp *kps_right, const uint8_t *desc_right, int n_right,
                                int mono_right, const float *level_sigma2, int nlevels, double ratio, int *left_to_right,
                                int *right_to_left, float *depth, float *p3d, int *counts) {
    if (!fisheye_args_ok(m, rig, kps_left, desc_left, n_left, mono_left, kps_right, desc_right, n_right, mono_right, level_sigma2, nlevels,
                         left_to_right, right_to_left, depth, p3d))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i = mono_left; i < n_left; i++)
        if (kps_left[i].octave < 0 || kps_left[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "keypoint octave outside the level table");
    for (int i = mono_right; i < n_right; i++)
        if (kps_right[i].octave < 0 || kps_right[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "keypoint octave outside the level table");
    int st;
    if ((st = upload(m, m->b[1], desc_left, (size_t)n_left * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[2], desc_right, (size_t)n_right * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[5], kps_left, (size_t)n_left * sizeof(orbx_kp))) != ORBX_OK) return st;
    if ((st = upload(m, m->b[6], kps_right, (size_t)n_right * sizeof(orbx_kp))) != ORBX_OK) return st;
    return fisheye_core(m, rig, m->b[5].as<orbx_kp>(), m->b[1].as<uint8_t>(), n_left, mono_left, m->b[6].as<orbx_kp>(), m->b[2].as<uint8_t>(),
                        n_right, mono_right, level_sigma2, nlevels, ratio, left_to_right, right_to_left, depth, p3d, counts);
}

int orbx_stereo_fisheye_matches_device(orbx_matcher *m, const orbx_kb8_stereo *rig, const orbx_kp *d_kps_left, const uint8_t *d_desc_left,
                                       int n_left, int mono_left, const orbx_kp *d_kps_right, const uint8_t *d_desc_right, int n_right,
                                       int mono_right, const float *level_sigma2, int nlevels, double ratio, int *left_to_right,
                                       int *right_to_left, float *depth, float *p3d, int *counts) {
    if (!fisheye_args_ok(m, rig, d_kps_left, d_desc_left, n_left, mono_left, d_kps_right, d_desc_right, n_right, mono_right, level_sigma2,
                         nlevels, left_to_right, right_to_left, depth, p3d))
        return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    return fisheye_core(m, rig, d_kps_left, d_desc_left, n_left, mono_left, d_kps_right, d_desc_right, n_right, mono_right, level_sigma2,
                        nlevels, ratio, left_to_right, right_to_left, depth, p3d, counts);
}

int orbx_distinctive_descriptors(orbx_matcher *m, const uint8_t *desc, const int *ptr, int ngroups, int *best) {
    if (!m || !ptr || !best || ngroups < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (ngroups == 0) return ORBX_OK;
    const int total = ptr[ngroups];
    if (total < 0 || (total && !desc)) return fail(ORBX_ERR_INVALID, "bad group table");
    for (int g = 0; g < ngroups; g++)
        if (ptr[g + 1] < ptr[g]) return fail(ORBX_ERR_INVALID, "group offsets must be non-decreasing");
    int st;
    if ((st = upload(m, m->b[1], desc, (size_t)total * 32)) != ORBX_OK) return st;
    if ((st = upload(m, m->b[5], ptr, (size_t)(ngroups + 1) * sizeof(int))) != ORBX_OK) return st;
    M_CUDA(m->b[7].ensure((size_t)ngroups * sizeof(int)));
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_distinctive<<<(ngroups + kDistWarps - 1) / kDistWarps, kDistWarps * 32, 0, m->stream>>>((const uint4 *)m->b[1].p, m->b[5].as<int>(), ngroups,
                                                                                            m->b[7].as<int>());
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(best, m->b[7].p, (size_t)ngroups * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_vocabulary_create(orbx_vocabulary **out, int device, int n_nodes, const int *child_ptr, const int *child, const uint8_t *node_desc,
                           const int *word_id, const double *weight, int levels) {
    if (!out || n_nodes < 1 || !child_ptr || !child || !node_desc || !word_id || !weight || levels < 1)
        return fail(ORBX_ERR_INVALID, "bad vocabulary arguments");
    *out = nullptr;
    if (child_ptr[0] != 0) return fail(ORBX_ERR_INVALID, "child_ptr[0] must be 0");
    for (int i = 0; i < n_nodes; i++)
        if (child_ptr[i + 1] < child_ptr[i]) return fail(ORBX_ERR_INVALID, "child_ptr must be non-decreasing");
    const int nch = child_ptr[n_nodes];
    for (int k = 0; k < nch; k++)
        if (child[k] <= 0 || child[k] >= n_nodes) return fail(ORBX_ERR_INVALID, "child id out of range");
    if (child_ptr[1] == 0) return fail(ORBX_ERR_INVALID, "the root has no children");
    const int ndev = orbx_device_count();
    if (ndev <= 0) return fail(ORBX_ERR_NO_DEVICE, "no CUDA device (liborbx has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ORBX_ERR_NO_DEVICE, "device index out of range");
    M_CUDA(cudaSetDevice(device));
    orbx_vocabulary *v = new orbx_vocabulary();
    v->device = device; v->n_nodes = n_nodes; v->n_children = nch; v->levels = levels;
    v->weight.assign(weight, weight + n_nodes); v->word.assign(word_id, word_id + n_nodes);
    {   // word id -> leaf node: owned by the vocabulary (a per-thread cache keyed by the handle's address would go stale when a
        // destroyed vocabulary's address is reused by the next one)
        int maxw = -1;
        for (int i = 0; i < n_nodes; i++) maxw = std::max(maxw, word_id[i]);
        v->leaf_of_word.assign((size_t)maxw + 1, -1);
        for (int i = 0; i < n_nodes; i++) if (word_id[i] >= 0) v->leaf_of_word[word_id[i]] = i;
    }
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_child_ptr, (size_t)(n_nodes + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_child, (size_t)std::max(nch, 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_word, (size_t)n_nodes * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_desc, (size_t)n_nodes * 32);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_ptr, child_ptr, (size_t)(n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child, child, (size_t)nch * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_word, word_id, (size_t)n_nodes * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { orbx_vocabulary_destroy(v); return fail(ORBX_ERR_CUDA, std::string("vocabulary upload: ") + cudaGetErrorString(e)); }
    *out = v;
    return ORBX_OK;
}

void orbx_vocabulary_destroy(orbx_vocabulary *v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_child_ptr); cudaFree(v->d_child); cudaFree(v->d_word); cudaFree(v->d_desc);
    delete v;
}

int orbx_vocabulary_transform(orbx_matcher *m, const orbx_vocabulary *v, const uint8_t *desc, int n, int levelsup, int *word_id,
                              int *node_id, double *weight) {
    if (!m || !v || n < 0 || !word_id || !node_id || (n && !desc)) return fail(ORBX_ERR_INVALID, "bad arguments");
    if (m->device != v->device) return fail(ORBX_ERR_INVALID, "matcher and vocabulary live on different devices");
    M_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (n == 0) return ORBX_OK;
    int st;
    if ((st = upload(m, m->b[1], desc, (size_t)n * 32)) != ORBX_OK) return st;
    M_CUDA(m->b[7].ensure((size_t)n * 2 * sizeof(int)));
    int *o = m->b[7].as<int>();
    M_CUDA(cudaEventRecord(m->ev0, m->stream));
    k_vocab_transform<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, m->stream>>>(v->d_child_ptr, v->d_child, (const uint4 *)v->d_desc, v->d_word,
                                                                                      v->levels - levelsup, (const uint4 *)m->b[1].p, n, o, o + n);
    m->last_launches = 1;
    M_CUDA(cudaGetLastError());
    M_CUDA(cudaEventRecord(m->ev1, m->stream));
    m->timed = true;
    M_CUDA(cudaMemcpyAsync(word_id, o, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaMemcpyAsync(node_id, o + n, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    M_CUDA(cudaStreamSynchronize(m->stream));
    if (weight) {   // word id -> weight of the leaf (idf for TF_IDF vocabularies); a leaf's word id indexes m_words, keep a reverse map
        const std::vector<int> &leaf_of_word = v->leaf_of_word;
        for (int i = 0; i < n; i++) weight[i] = (word_id[i] >= 0 && word_id[i] < (int)leaf_of_word.size() && leaf_of_word[word_id[i]] >= 0)
                                                    ? v->weight[leaf_of_word[word_id[i]]] : 0.0;
    }
    return ORBX_OK;
}

int orbx_popc_peak(int device, double *gpopc_per_s) {
    if (!gpopc_per_s) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (orbx_device_count() <= device || device < 0) return fail(ORBX_ERR_NO_DEVICE, "no such CUDA device");
    M_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint32_t *d = nullptr;
    M_CUDA(cudaMalloc((void **)&d, 64));
    cudaEvent_t e0, e1;
    M_CUDA(cudaEventCreate(&e0)); M_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 8;
    k_popc_peak<<<blocks, 256>>>(d, 64, 1);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        M_CUDA(cudaEventRecord(e0));
        k_popc_peak<<<blocks, 256>>>(d, iters, 1 + rep);
        M_CUDA(cudaEventRecord(e1));
        M_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        M_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    M_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *gpopc_per_s = (double)blocks * 256 * iters * 8 / (best * 1e-3) * 1e-9;
    return ORBX_OK;
}


// ---- guided matchers batched across frames: a pool of lanes (matcher handle + host worker thread each) ---------------------------
}  // extern "C"

struct orbx_matcher_pool {
    int device = 0;
    std::vector<orbx_matcher *> lanes;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<void(orbx_matcher *, int)> job;      // job(lane handle, call index)
    int n_calls = 0, generation = 0, running = 0;
    std::atomic<int> next{0};
    bool stop = false;
};

static void pool_worker(orbx_matcher_pool *p, int lane) {
    int seen = 0;
    cudaSetDevice(p->device);
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_work.wait(lk, [&] { return p->stop || p->generation != seen; });
            if (p->stop) return;
            seen = p->generation;
        }
        // a batch smaller than the pool uses the first lanes only: their buffers have their size from earlier calls (a lane's first
        // call allocates), so a one-frame batch costs what a plain call costs
        if (lane < p->n_calls)
            for (int i = p->next.fetch_add(1); i < p->n_calls; i = p->next.fetch_add(1)) p->job(p->lanes[lane], i);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            if (--p->running == 0) p->cv_done.notify_all();
        }
    }
}

static int pool_run(orbx_matcher_pool *p, int n, std::function<void(orbx_matcher *, int)> job) {
    if (!p || n < 0) return fail(ORBX_ERR_INVALID, "bad arguments");
    if (n == 0) return ORBX_OK;
    std::unique_lock<std::mutex> lk(p->mu);
    p->job = std::move(job); p->n_calls = n; p->next = 0; p->running = (int)p->workers.size(); p->generation++;
    p->cv_work.notify_all();
    p->cv_done.wait(lk, [&] { return p->running == 0; });
    return ORBX_OK;
}

extern "C" {

int orbx_matcher_pool_create(orbx_matcher_pool **out, int device, int lanes) {
    if (!out || lanes < 1 || lanes > 64) return fail(ORBX_ERR_INVALID, "bad arguments");
    *out = nullptr;
    orbx_matcher_pool *p = new orbx_matcher_pool();
    p->device = device;
    for (int l = 0; l < lanes; l++) {
        orbx_matcher *m = nullptr;
        const int st = orbx_matcher_create(&m, device);
        if (st != ORBX_OK) { for (auto *x : p->lanes) orbx_matcher_destroy(x); delete p; return st; }
        p->lanes.push_back(m);
    }
    for (int l = 0; l < lanes; l++) p->workers.emplace_back(pool_worker, p, l);
    *out = p;
    return ORBX_OK;
}

void orbx_matcher_pool_destroy(orbx_matcher_pool *p) {
    if (!p) return;
    { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; }
    p->cv_work.notify_all();
    for (auto &t : p->workers) t.join();
    for (auto *m : p->lanes) orbx_matcher_destroy(m);
    delete p;
}

int orbx_search_by_projection_batch(orbx_matcher_pool *p, int ncalls, orbx_sbp_call *c, int *status) {
    if (ncalls > 0 && (!c || !status)) return fail(ORBX_ERR_INVALID, "NULL argument");
    return pool_run(p, ncalls, [c, status](orbx_matcher *m, int i) {
        orbx_sbp_call &a = c[i];
        status[i] = orbx_search_by_projection(m, a.ncam, a.cam_off, a.kps, a.desc, a.grid_bounds, a.scale_factors, a.nlevels, a.occupied, a.l2r, a.r2l,
                                              a.u_right, a.nmp, a.in_view, a.proj_x, a.proj_y, a.level, a.view_cos, a.proj_xr, a.depth, a.bad, a.has_obs,
                                              a.mp_desc, a.th, a.far_points, a.th_far_points, a.nnratio, a.assigned, &a.nmatches);
    });
}

int orbx_search_by_projection_last_batch(orbx_matcher_pool *p, int ncalls, orbx_sbp_last_call *c, int *status) {
    if (ncalls > 0 && (!c || !status)) return fail(ORBX_ERR_INVALID, "NULL argument");
    return pool_run(p, ncalls, [c, status](orbx_matcher *m, int i) {
        orbx_sbp_last_call &a = c[i];
        status[i] = orbx_search_by_projection_last(m, a.ncam, a.cam_off, a.kps, a.desc, a.grid_bounds, a.scale_factors, a.nlevels, a.occupied, a.u_right,
                                                   a.npts, a.valid, a.proj_x, a.proj_y, a.proj_ur, a.last_octave, a.last_angle, a.has_obs, a.pt_desc, a.th,
                                                   a.level_gate, a.check_orientation, a.max_dist, a.assigned, &a.nmatches);
    });
}

int orbx_search_by_bow_frame_batch(orbx_matcher_pool *p, int ncalls, orbx_bow_call *c, int *status) {
    if (ncalls > 0 && (!c || !status)) return fail(ORBX_ERR_INVALID, "NULL argument");
    return pool_run(p, ncalls, [c, status](orbx_matcher *m, int i) {
        orbx_bow_call &a = c[i];
        status[i] = orbx_search_by_bow_frame(m, a.nn_kf, a.nodes_kf, a.ptr_kf, a.idx_kf, a.nn_f, a.nodes_f, a.ptr_f, a.idx_f, a.desc_kf, a.n_kf, a.desc_f,
                                             a.n_f, a.skip_kf, a.angle_kf, a.angle_f, a.ncam_f, a.cam_off_f, a.nnratio, a.check_orientation, a.match_f,
                                             &a.nmatches);
    });
}

}  // extern "C"
