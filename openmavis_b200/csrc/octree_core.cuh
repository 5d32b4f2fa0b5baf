// octree_core.cuh — DistributeOctTree (src/ORBextractor.cc:425-702) as a data-parallel array algorithm.
//
// The reference grows a std::list of quadtree nodes, copying each node's keypoints into its children.
// Here a node never owns keypoints: every candidate carries the list position of the node it currently
// falls in (`node[i]`), child membership is a pure function of the candidate's coordinates, node sizes are
// histogram counts, and the per-node winner ("first keypoint with the strictly greatest response",
// :683-699) is an atomicMax over (response, -candidate order).  Only the list ORDER is reproduced
// sequentially, with prefix sums:
//   full pass  (:551-613)  new list = reverse(non-empty children in visit order) ++ untouched bNoMore nodes
//   careful    (:619-680)  candidates sorted with an exact emulation of libstdc++ std::sort (the comparator
//                          leaves ties, so the introsort's permutation is part of the result), split from
//                          the back until the list reaches N.
// The code is written against a tiny execution context (thread id / count, barrier, atomics, block scan) so
// the very same source runs as one CUDA block per (image, level) and as a single host thread in the CPU
// test-suite (tests/test_host_core.py), where it is compared against the std::list-faithful oracle.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define OCT_HD __host__ __device__ __forceinline__
#else
#define OCT_HD inline
#endif

namespace orbx {

// ---- exact libstdc++ std::sort (introsort, threshold 16) on packed entries --------------------------------
// entry = key << 32 | payload; key = node size << 13 | UL.x  (compareNodes, :482-494: size, then UL.x).  The key is the
// high word, so a comparison is one 32-bit compare of the upper registers.
OCT_HD bool oct_lt(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
OCT_HD void oct_swap(uint64_t *a, int i, int j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }

OCT_HD void oct_push_heap(uint64_t *a, int first, int hole, int top, uint64_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && oct_lt(a[first + parent], value)) {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}
OCT_HD void oct_adjust_heap(uint64_t *a, int first, int hole, int len, uint64_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (oct_lt(a[first + child], a[first + child - 1])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + child - 1];
        hole = child - 1;
    }
    oct_push_heap(a, first, hole, top, value);
}
OCT_HD void oct_heapsort(uint64_t *a, int first, int last) {   // std::__partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            uint64_t v = a[first + parent];
            oct_adjust_heap(a, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) {
        --l;
        uint64_t v = a[l];
        a[l] = a[first];
        oct_adjust_heap(a, first, 0, l - first, v);
    }
}
// The sequential scans below read 4 elements ahead of the comparison that decides whether they are needed: the loads do
// not depend on each other, so one shared-memory latency covers 4 steps.  The array must be readable 3 entries beyond
// both ends (kOctSortPad); the values read there are never used.
constexpr int kOctSortPad = 4;

OCT_HD void oct_unguarded_linear_insert(uint64_t *a, int last) {
    const uint64_t val = a[last];
    int next = last - 1;
    for (;;) {
        const uint64_t x0 = a[next], x1 = a[next - 1], x2 = a[next - 2], x3 = a[next - 3];
        if (!oct_lt(val, x0)) break;
        a[last] = x0; last = next; --next;
        if (!oct_lt(val, x1)) break;
        a[last] = x1; last = next; --next;
        if (!oct_lt(val, x2)) break;
        a[last] = x2; last = next; --next;
        if (!oct_lt(val, x3)) break;
        a[last] = x3; last = next; --next;
    }
    a[last] = val;
}
// first f' >= f with !(a[f'] < piv)   (the caller guarantees a sentinel inside the range)
OCT_HD int oct_scan_up(const uint64_t *a, int f, uint64_t piv) {
    for (;;) {
        const uint64_t x0 = a[f], x1 = a[f + 1], x2 = a[f + 2], x3 = a[f + 3];
        if (!oct_lt(x0, piv)) return f;
        if (!oct_lt(x1, piv)) return f + 1;
        if (!oct_lt(x2, piv)) return f + 2;
        if (!oct_lt(x3, piv)) return f + 3;
        f += 4;
    }
}
// first l' <= l with !(piv < a[l'])
OCT_HD int oct_scan_down(const uint64_t *a, int l, uint64_t piv) {
    for (;;) {
        const uint64_t x0 = a[l], x1 = a[l - 1], x2 = a[l - 2], x3 = a[l - 3];
        if (!oct_lt(piv, x0)) return l;
        if (!oct_lt(piv, x1)) return l - 1;
        if (!oct_lt(piv, x2)) return l - 2;
        if (!oct_lt(piv, x3)) return l - 3;
        l -= 4;
    }
}
OCT_HD void oct_insertion_sort(uint64_t *a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (oct_lt(a[i], a[first])) {
            uint64_t val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        } else {
            oct_unguarded_linear_insert(a, i);
        }
    }
}
// stack: at most 2*floor(log2 n)+1 pending ranges
OCT_HD void oct_std_sort(uint64_t *a, int n) {
    if (n <= 1) return;
    int lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { oct_heapsort(a, first, last); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first+1, mid, last-1)
                const int ia = first + 1, ib = mid, ic = last - 1;
                if (oct_lt(a[ia], a[ib])) {
                    if (oct_lt(a[ib], a[ic])) oct_swap(a, first, ib);
                    else if (oct_lt(a[ia], a[ic])) oct_swap(a, first, ic);
                    else oct_swap(a, first, ia);
                } else if (oct_lt(a[ia], a[ic])) oct_swap(a, first, ia);
                else if (oct_lt(a[ib], a[ic])) oct_swap(a, first, ic);
                else oct_swap(a, first, ib);
            }
            int f = first + 1, l = last;   // __unguarded_partition(first+1, last, pivot=first)
            for (;;) {
                const uint64_t piv = a[first];      // the pivot slot is never swapped inside the loop (f > first)
                f = oct_scan_up(a, f, piv);
                l = oct_scan_down(a, l - 1, piv);
                if (!(f < l)) break;
                oct_swap(a, f, l);
                ++f;
            }
            const int cut = f;
            // recurse on [cut, last) first (the reference recursion), continue the loop on [first, cut):
            // both ranges are disjoint, so the order of processing does not change the result.
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    if (n > 16) {
        oct_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) oct_unguarded_linear_insert(a, i);
    } else {
        oct_insertion_sort(a, 0, n);
    }
}

// ---- the same std::sort, restructured for a thread block -----------------------------------------------------------------
// libstdc++'s sort = introsort loop (partitions down to ranges of <= 16) + one final insertion sort.  Two facts make it parallel
// without changing a single comparison result:
//  * a partition leaves [first, cut) <= pivot <= [cut, last), and the two ranges are processed independently afterwards, so all
//    ranges of one recursion depth can be partitioned at the same time, one thread each (breadth-first instead of depth-first);
//  * the final insertion sort never moves an element across a range boundary (val < a[boundary - 1] is false), and an insertion
//    sort is a STABLE sort: the final pass is a stable sort of every leaf range, i.e. a rank computation per element
//    (#smaller keys + #equal keys in front of it), fully parallel.  A range that hit the depth limit is heap-sorted by its thread
//    (already ordered -> the final pass leaves it alone).
// Scratch: rf/rl/rd = two range lists (first, last, depth) of `cap` entries each, lohi[cap] = leaf bounds per element,
// tmp[cap] = output of the rank pass, cnt[2] = list / work counters.  Every thread of the block calls this.
struct OctSortScratch { int *rf[2], *rl[2], *rd[2]; int *lohi; uint64_t *tmp; int *cnt; };

OCT_HD int oct_partition_range(uint64_t *a, int first, int last) {   // __unguarded_partition_pivot(first, last)
    const int mid = first + (last - first) / 2;
    {   // __move_median_to_first(first, first+1, mid, last-1)
        const int ia = first + 1, ib = mid, ic = last - 1;
        if (oct_lt(a[ia], a[ib])) {
            if (oct_lt(a[ib], a[ic])) oct_swap(a, first, ib);
            else if (oct_lt(a[ia], a[ic])) oct_swap(a, first, ic);
            else oct_swap(a, first, ia);
        } else if (oct_lt(a[ia], a[ic])) oct_swap(a, first, ia);
        else if (oct_lt(a[ib], a[ic])) oct_swap(a, first, ic);
        else oct_swap(a, first, ib);
    }
    int f = first + 1, l = last;   // __unguarded_partition(first+1, last, pivot=first)
    const uint64_t piv = a[first];  // the pivot slot is never swapped inside the loop (f > first)
    for (;;) {
        f = oct_scan_up(a, f, piv);
        l = oct_scan_down(a, l - 1, piv);
        if (!(f < l)) break;
        oct_swap(a, f, l);
        ++f;
    }
    return f;
}

template <class Ctx>
OCT_HD void oct_std_sort_block(Ctx &cx, uint64_t *a, int n, OctSortScratch &S) {
    if (n <= 1) return;
    if (cx.tid == 0) {
        int lg = 0;
        while ((n >> (lg + 1)) > 0) lg++;
        S.rf[0][0] = 0; S.rl[0][0] = n; S.rd[0][0] = 2 * lg;
        S.cnt[0] = 1; S.cnt[1] = 0;
    }
    cx.sync();
    int cur = 0;
    for (;;) {
        const int nr = S.cnt[cur];
        if (nr == 0) break;
        cx.sync();                                   // everybody has read the count before it is reset below
        if (cx.tid == 0) S.cnt[cur ^ 1] = 0;
        cx.sync();
        for (int r = cx.tid; r < nr; r += cx.nt) {
            const int first = S.rf[cur][r], last = S.rl[cur][r];
            int depth = S.rd[cur][r];
            if (last - first > 16 && depth > 0) {
                --depth;
                const int cut = oct_partition_range(a, first, last);
                const int o = cx.atomic_add(&S.cnt[cur ^ 1], 2);
                S.rf[cur ^ 1][o] = cut; S.rl[cur ^ 1][o] = last; S.rd[cur ^ 1][o] = depth;
                S.rf[cur ^ 1][o + 1] = first; S.rl[cur ^ 1][o + 1] = cut; S.rd[cur ^ 1][o + 1] = depth;
            } else if (last - first > 16) {          // depth limit: std::__partial_sort(first, last, last)
                oct_heapsort(a, first, last);
                for (int e = first; e < last; e++) S.lohi[e] = e | ((e + 1) << 16);
            } else {                                 // leaf: the final insertion sort orders it stably
                for (int e = first; e < last; e++) S.lohi[e] = first | (last << 16);
            }
        }
        cx.sync();
        cur ^= 1;
    }
    cx.sync();
    for (int e = cx.tid; e < n; e += cx.nt) {        // stable rank inside the leaf range
        const int lo = S.lohi[e] & 0xFFFF, hi = S.lohi[e] >> 16;
        const uint64_t v = a[e];
        int rank = 0;
        for (int j = lo; j < hi; j++) {
            const uint64_t u = a[j];
            rank += (oct_lt(u, v) || (j < e && !oct_lt(v, u))) ? 1 : 0;
        }
        S.tmp[lo + rank] = v;
    }
    cx.sync();
    for (int e = cx.tid; e < n; e += cx.nt) a[e] = S.tmp[e];
    cx.sync();
}

// ---- node tables (one block's shared memory, or host arrays) ----------------------------------------------
struct OctBox { short x0, x1, y0, y1; };   // UL.x, UR.x, UL.y, BR.y

struct OctTable {
    OctBox *box;
    int *cnt;
    unsigned char *nomore;
};

struct OctWork {
    int cap;                 // max list length
    OctTable t[2];
    int *child_cnt;          // [4*cap]
    int *child_pos;          // [4*cap]
    int *scan_a;             // [cap+1]
    int *scan_b;             // [cap+1]
    int *scan_c;             // [cap+1]
    int *remap;              // [cap]
    int *clist;              // [cap]   candidate nodes (list positions) in creation order
    int *cidx;               // [cap]   list position -> index in clist, or -1
    int *order;              // [cap]   careful phase: clist indices in processing order
    uint64_t *srt;           // [cap]
    unsigned *best;          // [cap]
    int *vars;               // [8] block-shared scalars
};

OCT_HD size_t oct_work_bytes(int cap) {
    size_t b = 0;
    b += 2 * (size_t)cap * (sizeof(OctBox) + sizeof(int)) + 2 * (((size_t)cap + 7) & ~(size_t)7);
    b += (size_t)cap * 4 * sizeof(int) * 2;           // child_cnt, child_pos
    b += 3 * ((size_t)cap + 1) * sizeof(int);          // scans
    b += 4 * (size_t)cap * sizeof(int);                // remap, clist, cidx, order
    b += ((size_t)cap + 2 * kOctSortPad) * sizeof(uint64_t);   // srt (+ read-ahead pads)
    b += (size_t)cap * sizeof(unsigned);               // best
    b += 16 * sizeof(int);
    return b + 64;
}

OCT_HD void oct_carve(OctWork &w, void *mem, int cap) {
    unsigned char *p = (unsigned char *)mem;
    p = (unsigned char *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    w.cap = cap;
    w.srt = (uint64_t *)p + kOctSortPad; p += ((size_t)cap + 2 * kOctSortPad) * sizeof(uint64_t);
    for (int k = 0; k < 2; k++) { w.t[k].box = (OctBox *)p; p += (size_t)cap * sizeof(OctBox); }
    for (int k = 0; k < 2; k++) { w.t[k].cnt = (int *)p; p += (size_t)cap * sizeof(int); }
    w.child_cnt = (int *)p; p += (size_t)cap * 4 * sizeof(int);
    w.child_pos = (int *)p; p += (size_t)cap * 4 * sizeof(int);
    w.scan_a = (int *)p; p += ((size_t)cap + 1) * sizeof(int);
    w.scan_b = (int *)p; p += ((size_t)cap + 1) * sizeof(int);
    w.scan_c = (int *)p; p += ((size_t)cap + 1) * sizeof(int);
    w.remap = (int *)p; p += (size_t)cap * sizeof(int);
    w.clist = (int *)p; p += (size_t)cap * sizeof(int);
    w.cidx = (int *)p; p += (size_t)cap * sizeof(int);
    w.order = (int *)p; p += (size_t)cap * sizeof(int);
    w.best = (unsigned *)p; p += (size_t)cap * sizeof(unsigned);
    w.vars = (int *)p; p += 16 * sizeof(int);
    for (int k = 0; k < 2; k++) { w.t[k].nomore = p; p += ((size_t)cap + 7) & ~(size_t)7; }
}

// candidate packing: x | y << 12 | score << 24
OCT_HD int cand_x(uint32_t c) { return (int)(c & 0xFFFu); }
OCT_HD int cand_y(uint32_t c) { return (int)((c >> 12) & 0xFFFu); }
OCT_HD int cand_score(uint32_t c) { return (int)(c >> 24); }
OCT_HD uint32_t cand_pack(int x, int y, int score) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)score << 24); }

OCT_HD int oct_quadrant(const OctBox &b, uint32_t c) {
    // DivideNode (:425-480): halfX = ceil((UR.x-UL.x)/2), halfY = ceil((BR.y-UL.y)/2);
    // n1: x < UL.x+halfX && y < UL.y+halfY ; n2: x >= .. && y < .. ; n3: x < .. && y >= .. ; n4: rest
    const int mx = b.x0 + ((b.x1 - b.x0 + 1) >> 1);
    const int my = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
    return (cand_x(c) >= mx ? 1 : 0) + (cand_y(c) >= my ? 2 : 0);
}
OCT_HD OctBox oct_child_box(const OctBox &b, int q) {
    const short mx = (short)(b.x0 + ((b.x1 - b.x0 + 1) >> 1));
    const short my = (short)(b.y0 + ((b.y1 - b.y0 + 1) >> 1));
    OctBox c;
    c.x0 = (q & 1) ? mx : b.x0; c.x1 = (q & 1) ? b.x1 : mx;
    c.y0 = (q & 2) ? my : b.y0; c.y1 = (q & 2) ? b.y1 : my;
    return c;
}

// Ctx requirements: int tid, nt; void sync(); int atomic_add(int*, int); void atomic_max(unsigned*, unsigned);
//                   void hist_add(int *cnt, int key, bool valid, int nkeys) — ++cnt[key] for the valid callers; called by EVERY thread of
//                   the block in lock step; with few keys (nkeys <= 32) most callers share a counter: aggregate per warp first
//                   int excl_scan(int *arr, int n)  — block-wide, in place, returns the total to every thread
//                   (callers bracket it with sync()).
//
// cand[n]  : candidates in vToDistributeKeys order (border-relative coordinates)
// node[n]  : scratch, one entry per candidate
// sel[cap] : output, candidate indices in final list order.  Returns the number of selected keypoints.
template <class Ctx>
OCT_HD int oct_distribute(Ctx &cx, OctWork &w, const uint32_t *cand, unsigned short *node, int n, int x_span,
                          int y_span, int n_ini, float h_x, int N, int *sel) {
    if (n <= 0) return 0;
    int cur = 0;
    int *V = w.vars;   // V[0]=list size, V[1]=#candidate nodes, V[2]=state (0 full pass, 1 careful, 2 done), V[3]=prev size
    // ---- roots (:505-542)
    for (int r = cx.tid; r < n_ini; r += cx.nt) w.child_cnt[r] = 0;
    cx.sync();
    // (all sweeps over the candidates: uniform trip count — hist_add is a whole-block call — and 4 independent loads in flight)
    for (int base = 0; base < n; base += 4 * cx.nt) {
        uint32_t cd[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = base + u * cx.nt + cx.tid; cd[u] = i < n ? cand[i] : 0u; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * cx.nt + cx.tid;
            const bool v = i < n;
            const int r = v ? (int)((float)cand_x(cd[u]) / h_x) : 0;
            if (v) node[i] = (unsigned short)r;
            cx.hist_add(w.child_cnt, r, v, n_ini);          // a handful of counters shared by every candidate
        }
    }
    cx.sync();
    if (cx.tid == 0) {
        int ls = 0;
        for (int r = 0; r < n_ini; r++) {
            w.remap[r] = -1;
            if (w.child_cnt[r] == 0) continue;
            OctBox b;
            b.x0 = (short)(int)(h_x * (float)r); b.x1 = (short)(int)(h_x * (float)(r + 1));
            b.y0 = 0; b.y1 = (short)y_span;
            w.t[cur].box[ls] = b; w.t[cur].cnt[ls] = w.child_cnt[r];
            w.t[cur].nomore[ls] = w.child_cnt[r] == 1;
            w.remap[r] = ls++;
        }
        V[0] = ls; V[1] = 0; V[2] = 0; V[3] = -1;
    }
    cx.sync();
    // Labels: node[i] is a KEY, the candidate's list position is map[node[i]] with map = w.child_pos.  Every round stores the key
    // 4 * position + quadrant while it counts, and publishes the map of those keys when the new list is known — one sweep over
    // the candidates per round instead of two.
    for (int r = cx.tid; r < n_ini; r += cx.nt) w.child_pos[r] = w.remap[r];
    cx.sync();
    (void)x_span;

    for (;;) {
        const int state = V[2];
        if (state == 2) break;
        const int ls = V[0];
        OctTable &T = w.t[cur];
        OctTable &U = w.t[cur ^ 1];
        if (state == 0) {
            // ------------------------------------------------ full pass (:551-613)
            for (int k = cx.tid; k < 4 * ls; k += cx.nt) w.child_cnt[k] = 0;
            cx.sync();
            for (int base = 0; base < n; base += 4 * cx.nt) {
                uint32_t cd[4]; int pn[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + u * cx.nt + cx.tid;
                    cd[u] = i < n ? cand[i] : 0u; pn[u] = i < n ? w.child_pos[node[i]] : -1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int p = pn[u];
                    const bool v = p >= 0 && !T.nomore[p];
                    const int key = v ? 4 * p + oct_quadrant(T.box[p], cd[u]) : (p > 0 ? 4 * p : 0);
                    if (p >= 0) node[base + u * cx.nt + cx.tid] = (unsigned short)key;
                    cx.hist_add(w.child_cnt, key, v, 4 * ls);
                }
            }
            cx.sync();
            for (int p = cx.tid; p < ls; p += cx.nt) {
                int a = 0, e = 0;
                if (!T.nomore[p])
                    for (int q = 0; q < 4; q++) { const int c = w.child_cnt[4 * p + q]; a += c > 0; e += c > 1; }
                w.scan_a[p] = a; w.scan_b[p] = T.nomore[p] ? 1 : 0; w.scan_c[p] = e;
            }
            cx.sync();
            const int A = cx.excl_scan(w.scan_a, ls);
            cx.sync();
            const int B = cx.excl_scan(w.scan_b, ls);
            cx.sync();
            const int E = cx.excl_scan(w.scan_c, ls);
            cx.sync();
            for (int p = cx.tid; p < ls; p += cx.nt) {
                if (T.nomore[p]) {
                    const int np = A + w.scan_b[p];
                    w.child_pos[4 * p] = np;
                    U.box[np] = T.box[p]; U.cnt[np] = T.cnt[p]; U.nomore[np] = 1;
                } else {
                    int j = 0, jj = 0;
                    for (int q = 0; q < 4; q++) {
                        const int c = w.child_cnt[4 * p + q];
                        if (c == 0) continue;
                        const int np = A - 1 - (w.scan_a[p] + j);
                        w.child_pos[4 * p + q] = np;
                        U.box[np] = oct_child_box(T.box[p], q); U.cnt[np] = c; U.nomore[np] = c == 1;
                        if (c > 1) { w.clist[w.scan_c[p] + jj] = np; jj++; }
                        j++;
                    }
                }
            }
            cx.sync();
            if (cx.tid == 0) {
                const int ns = A + B;
                V[1] = E;
                if (ns >= N || ns == ls) V[2] = 2;
                else if (ns + 3 * E > N) V[2] = 1;
                V[0] = ns; V[3] = ls;
            }
            cur ^= 1;
            cx.sync();
        } else {
            // ------------------------------------------------ careful round (:619-680)
            const int m = V[1];
            for (int p = cx.tid; p < ls; p += cx.nt) w.cidx[p] = -1;
            for (int k = cx.tid; k < 4 * m; k += cx.nt) w.child_cnt[k] = 0;
            cx.sync();
            for (int c = cx.tid; c < m; c += cx.nt) {
                const int p = w.clist[c];
                w.cidx[p] = c;
                w.srt[c] = ((uint64_t)(((uint32_t)T.cnt[p] << 13) | (uint32_t)T.box[p].x0) << 32) | (uint64_t)c;
            }
            cx.sync();
            for (int base = 0; base < n; base += 4 * cx.nt) {
                uint32_t cd[4]; int pn[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + u * cx.nt + cx.tid;
                    cd[u] = i < n ? cand[i] : 0u; pn[u] = i < n ? w.child_pos[node[i]] : -1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int p = pn[u];
                    const int c = p >= 0 ? w.cidx[p] : -1;
                    const bool v = c >= 0;
                    const int q = v ? oct_quadrant(T.box[p], cd[u]) : 0;
                    if (p >= 0) node[base + u * cx.nt + cx.tid] = (unsigned short)(4 * p + q);
                    cx.hist_add(w.child_cnt, 4 * c + q, v, 4 * m);
                }
            }
            cx.sync();   // (the old map is dead: the sort below uses its storage)
            {   // exact std::sort of the candidates, block-parallel; scratch = arrays that are dead until the tables are rebuilt
                OctSortScratch S;
                S.rf[0] = w.scan_a; S.rl[0] = w.scan_b; S.rd[0] = w.scan_c;
                S.rf[1] = w.remap; S.rl[1] = w.order; S.rd[1] = (int *)w.best;
                S.lohi = w.child_pos; S.tmp = (uint64_t *)(w.child_pos + 2 * ((w.cap + 1) / 2)); S.cnt = V + 8;
                oct_std_sort_block(cx, w.srt, m, S);
            }
            if (cx.tid == 0) {
                int size = ls, np = 0, A = 0, E = 0;
                for (int j = m - 1; j >= 0; j--) {
                    const int c = (int)(w.srt[j] & 0xFFFF);
                    int nne = 0, e = 0;
                    for (int q = 0; q < 4; q++) { const int cc = w.child_cnt[4 * c + q]; nne += cc > 0; e += cc > 1; }
                    w.order[np] = c; w.scan_a[np] = A; w.scan_c[np] = E;
                    A += nne; E += e; np++;
                    size += nne - 1;
                    if (size >= N) break;
                }
                V[4] = np; V[5] = A; V[6] = E; V[7] = size;
            }
            cx.sync();
            const int np = V[4], A = V[5], E = V[6], ns = V[7];
            // processed parents leave the list; everybody else keeps the old order behind the new children
            for (int p = cx.tid; p < ls; p += cx.nt) w.scan_b[p] = 1;
            cx.sync();
            for (int k = cx.tid; k < np; k += cx.nt) w.scan_b[w.clist[w.order[k]]] = 0;
            cx.sync();
            cx.excl_scan(w.scan_b, ls);
            cx.sync();
            for (int k = cx.tid; k < np; k += cx.nt) w.cidx[w.clist[w.order[k]]] = -2 - k;   // mark processed (k recoverable)
            cx.sync();
            for (int p = cx.tid; p < ls; p += cx.nt) {
                if (w.cidx[p] <= -2) continue;
                const int q = A + w.scan_b[p];
                w.child_pos[4 * p] = q; w.child_pos[4 * p + 1] = q; w.child_pos[4 * p + 2] = q; w.child_pos[4 * p + 3] = q;
                U.box[q] = T.box[p]; U.cnt[q] = T.cnt[p]; U.nomore[q] = T.nomore[p];
            }
            // new candidate list is written to scan-free storage: reuse `remap`-independent array w.best as temp
            for (int k = cx.tid; k < np; k += cx.nt) {
                const int c = w.order[k];
                const int p = w.clist[c];
                int j = 0, jj = 0;
                for (int q = 0; q < 4; q++) {
                    const int cc = w.child_cnt[4 * c + q];
                    if (cc == 0) continue;
                    const int q2 = A - 1 - (w.scan_a[k] + j);
                    w.child_pos[4 * p + q] = q2;
                    U.box[q2] = oct_child_box(T.box[p], q); U.cnt[q2] = cc; U.nomore[q2] = cc == 1;
                    if (cc > 1) { w.best[w.scan_c[k] + jj] = (unsigned)q2; jj++; }
                    j++;
                }
            }
            cx.sync();
            for (int k = cx.tid; k < E; k += cx.nt) w.clist[k] = (int)w.best[k];
            if (cx.tid == 0) {
                V[1] = E;
                if (ns >= N || ns == ls) V[2] = 2;
                V[0] = ns; V[3] = ls;
            }
            cur ^= 1;
            cx.sync();
        }
    }
    // ---- winner per node (:683-699)
    const int ls = V[0];
    for (int p = cx.tid; p < ls; p += cx.nt) w.best[p] = 0;
    cx.sync();
    for (int base = 0; base < n; base += 4 * cx.nt) {
        uint32_t cd[4]; int pn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * cx.nt + cx.tid;
            cd[u] = i < n ? cand[i] : 0u; pn[u] = i < n ? w.child_pos[node[i]] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (pn[u] >= 0) cx.atomic_max(&w.best[pn[u]], ((uint32_t)cand_score(cd[u]) << 24) | (uint32_t)(0xFFFFFF - (base + u * cx.nt + cx.tid)));
    }
    cx.sync();
    for (int p = cx.tid; p < ls; p += cx.nt) sel[p] = 0xFFFFFF - (int)(w.best[p] & 0xFFFFFF);
    cx.sync();
    return ls;
}

// Upper bound of the list length for target N with n_ini roots: the careful phase stops
// within +2 of N, a guarded full pass never exceeds N, the unguarded first pass yields at most 4*n_ini.
OCT_HD int oct_list_cap(int N, int n_ini) {
    int a = N + 3, b = 4 * n_ini;
    return (a > b ? a : b) + 1;
}

}  // namespace orbx
