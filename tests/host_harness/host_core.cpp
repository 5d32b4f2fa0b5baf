// Host harness: runs the product's host/device-shared cores (openmavis_b200/csrc/*_core.cuh) on ONE CPU thread
// so that their logic can be compared with the oracle in the CPU test-suite.  Not a product path.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../openmavis_b200/csrc/octree_core.cuh"

namespace {
struct HostCtx {
    int tid = 0, nt = 1;
    void sync() {}
    int atomic_add(int *p, int v) { int o = *p; *p += v; return o; }
    void atomic_max(unsigned *p, unsigned v) { if (v > *p) *p = v; }
    void hist_add(int *cnt, int key, bool valid, int) { if (valid) cnt[key]++; }
    int excl_scan(int *a, int n) { int s = 0; for (int i = 0; i < n; i++) { int v = a[i]; a[i] = s; s += v; } return s; }
};
}  // namespace

extern "C" {

void hc_std_sort(const int *sizes, const int *ulx, int n, int *perm) {
    std::vector<uint64_t> buf(n + 2 * orbx::kOctSortPad, 0);   // the sort reads (never uses) a few entries beyond both ends
    uint64_t *a = buf.data() + orbx::kOctSortPad;
    for (int i = 0; i < n; i++) a[i] = ((uint64_t)(((uint32_t)sizes[i] << 13) | (uint32_t)ulx[i]) << 32) | (uint64_t)i;
    orbx::oct_std_sort(a, n);
    for (int i = 0; i < n; i++) perm[i] = (int)(a[i] & 0xFFFF);
}

// the block-parallel formulation of the same sort (executed by one host thread: same comparisons, breadth-first)
void hc_std_sort_block(const int *sizes, const int *ulx, int n, int *perm) {
    std::vector<uint64_t> buf(n + 2 * orbx::kOctSortPad, 0), tmp(n + 1);
    uint64_t *a = buf.data() + orbx::kOctSortPad;
    for (int i = 0; i < n; i++) a[i] = ((uint64_t)(((uint32_t)sizes[i] << 13) | (uint32_t)ulx[i]) << 32) | (uint64_t)i;
    std::vector<int> r[6];
    for (auto &v : r) v.assign(n + 2, 0);
    std::vector<int> lohi(n + 2), cnt(4);
    orbx::OctSortScratch S;
    S.rf[0] = r[0].data(); S.rl[0] = r[1].data(); S.rd[0] = r[2].data();
    S.rf[1] = r[3].data(); S.rl[1] = r[4].data(); S.rd[1] = r[5].data();
    S.lohi = lohi.data(); S.tmp = tmp.data(); S.cnt = cnt.data();
    HostCtx cx;
    orbx::oct_std_sort_block(cx, a, n, S);
    for (int i = 0; i < n; i++) perm[i] = (int)(a[i] & 0xFFFF);
}

int hc_distribute(const int *xs, const int *ys, const int *scores, int n, int minX, int maxX, int minY, int maxY, int N,
                  int *sel, int cap) {
    using namespace orbx;
    const int n_ini = (int)roundf((float)(maxX - minX) / (maxY - minY));
    const float h_x = (float)(maxX - minX) / n_ini;
    const int lcap = oct_list_cap(N, n_ini);
    std::vector<unsigned char> mem(oct_work_bytes(lcap));
    OctWork w; oct_carve(w, mem.data(), lcap);
    std::vector<uint32_t> cand(n);
    for (int i = 0; i < n; i++) cand[i] = cand_pack(xs[i], ys[i], scores[i]);
    std::vector<unsigned short> node(n);
    std::vector<int> out(lcap);
    HostCtx cx;
    int k = oct_distribute(cx, w, cand.data(), node.data(), n, maxX - minX, maxY - minY, n_ini, h_x, N, out.data());
    for (int i = 0; i < k && i < cap; i++) sel[i] = out[i];
    return k;
}

}  // extern "C"
