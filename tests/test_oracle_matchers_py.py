"""CPU: a second, independent restatement (plain Python, written from the reference text) of the ordered matchers, compared with the
C++ oracle on small dense cases.  The reference holds no tests for these functions, so two independent readings
that agree are the pin for the oracle that the GPU parity tests then use:
  SearchByProjection(CurrentFrame, LastFrame)   src/ORBmatcher.cc:1985-2413 (+ the relocalisation / KeyFrame-Sim3 overloads)
  SearchByBoW(KeyFrame, Frame)                  :349-666
  SearchByBoW(KeyFrame, KeyFrame)               :1006-1129
  SearchForInitialization                       :895-1004
  Frame::AssignFeaturesToGrid / GetFeaturesInArea  src/Frame.cc:541-582, 890-978"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_descriptors

f32 = np.float32
COLS, ROWS = 64, 48


class PyGrid:
    def __init__(self, kps, bounds):
        self.kps = kps
        self.mn_x, self.mx_x, self.mn_y, self.mx_y = (f32(b) for b in bounds)
        self.inv_w = f32(COLS) / (self.mx_x - self.mn_x); self.inv_h = f32(ROWS) / (self.mx_y - self.mn_y)
        self.cells = [[[] for _ in range(ROWS)] for _ in range(COLS)]
        for i in range(kps.size):                                             # AssignFeaturesToGrid + PosInGrid (round())
            px = int(math.floor(float(f32((f32(kps['x'][i]) - self.mn_x) * self.inv_w)) + 0.5))
            py = int(math.floor(float(f32((f32(kps['y'][i]) - self.mn_y) * self.inv_h)) + 0.5))
            if 0 <= px < COLS and 0 <= py < ROWS:
                self.cells[px][py].append(i)

    def area(self, x, y, r, min_level, max_level):                            # GetFeaturesInArea
        x, y, r = f32(x), f32(y), f32(r)
        c0 = max(0, int(math.floor(float(f32(f32(f32(x - self.mn_x) - r) * self.inv_w)))))
        if c0 >= COLS:
            return []
        c1 = min(COLS - 1, int(math.ceil(float(f32(f32(f32(x - self.mn_x) + r) * self.inv_w)))))
        if c1 < 0:
            return []
        r0 = max(0, int(math.floor(float(f32(f32(f32(y - self.mn_y) - r) * self.inv_h)))))
        if r0 >= ROWS:
            return []
        r1 = min(ROWS - 1, int(math.ceil(float(f32(f32(f32(y - self.mn_y) + r) * self.inv_h)))))
        if r1 < 0:
            return []
        check = (min_level > 0) or (max_level >= 0)
        out = []
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for i in self.cells[ix][iy]:
                    o = int(self.kps['octave'][i])
                    if check and (o < min_level or (max_level >= 0 and o > max_level)):
                        continue
                    if abs(f32(self.kps['x'][i]) - x) < r and abs(f32(self.kps['y'][i]) - y) < r:
                        out.append(i)
        return out


def hamming(a, b):
    return int(np.unpackbits(a ^ b).sum())


def rot_bin(a1, a2):
    rot = f32(a1) - f32(a2)
    if rot < 0.0:
        rot = f32(rot + f32(360.0))
    v = float(f32(rot * f32(f32(1.0) / f32(30))))
    b = int(math.floor(v + 0.5))                                              # round(): half away from zero, v >= 0
    return 0 if b == 30 else b


def three_maxima_filter(hist):
    """ComputeThreeMaxima (:2537-2573) -> set of kept bins."""
    m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
    for i in range(30):
        s = len(hist[i])
        if s > m1:
            m3, m2, m1 = m2, m1, s; i3, i2, i1 = i2, i1, i
        elif s > m2:
            m3, m2 = m2, s; i3, i2 = i2, i
        elif s > m3:
            m3 = s; i3 = i
    if m2 < float(f32(0.1) * f32(m1)):
        i2 = i3 = -1
    elif m3 < float(f32(0.1) * f32(m1)):
        i3 = -1
    return {i1, i2, i3}


def random_kps(rng, n, size=120.0):
    k = np.zeros(n, O.KP_DTYPE)
    k['x'] = rng.uniform(0, size, n); k['y'] = rng.uniform(0, size, n); k['octave'] = rng.integers(0, 4, n)
    k['angle'] = rng.uniform(0, 360, n); k['class_id'] = -1
    return k


def pooled_desc(rng, n, pool):
    return pool[rng.integers(0, pool.shape[0], n)] ^ np.packbits(rng.random((n, 256)) < 0.03, axis=1, bitorder='little')


@pytest.mark.parametrize('seed,gate,check,max_dist', [(1, 0, True, 100), (2, 1, True, 100), (3, 2, False, 100), (4, 3, False, 40), (5, 0, True, 64)])
def test_search_by_projection_last_frame_two_restatements(seed, gate, check, max_dist):
    rng = np.random.default_rng(seed)
    ncam = 1 if gate == 3 else int(rng.choice([1, 2]))
    ns = [int(rng.integers(40, 90)) for _ in range(ncam)]
    cam_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    N = int(cam_off[-1])
    kps = random_kps(rng, N); pool = synth_descriptors(5, seed)
    desc = pooled_desc(rng, N, pool)
    sf = np.array([f32(1.2) ** l for l in range(8)], np.float32)
    occupied = (rng.random(N) < 0.1).astype(np.uint8)
    n = 150
    tgt = rng.integers(0, N, n)
    px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
    for c in range(ncam):
        kk = np.where((tgt >= cam_off[c]) & (tgt < cam_off[c + 1]), tgt, rng.integers(cam_off[c], cam_off[c + 1], n))
        px[:, c] = kps['x'][kk] + rng.uniform(-3, 3, n); py[:, c] = kps['y'][kk] + rng.uniform(-3, 3, n)
    empty0 = rng.random(n) < 0.1
    px[empty0, 0] = 5000                                                       # empty camera-0 window: the point's other cameras are skipped
    valid = (rng.random(n) < 0.9).astype(np.uint8)
    octv = np.clip(kps['octave'][tgt] + rng.integers(-1, 2, n), 0, 3).astype(np.int32)
    ang = ((kps['angle'][tgt] + np.where(rng.random(n) < 0.8, 25.0, rng.uniform(0, 360, n))) % 360).astype(np.float32)
    pdesc = pooled_desc(rng, n, pool)
    has_obs = (rng.random(n) < 0.8).astype(np.uint8)
    th = f32(4.0)
    # ---- Python restatement
    grids = [PyGrid(kps[cam_off[c]:cam_off[c + 1]], (0, 120, 0, 120)) for c in range(ncam)]
    occ = occupied.copy(); assigned = np.full(N, -1, np.int32); hist = [[] for _ in range(30)]; nm = 0
    for i in range(n):
        if not valid[i]:
            continue
        lvl = int(octv[i]); radius = f32(th * sf[lvl])
        for c in range(ncam):
            ck = kps[cam_off[c]:cam_off[c + 1]]
            if gate == 1:
                cand = grids[c].area(px[i, c], py[i, c], radius, lvl, -1)
            elif gate == 2:
                cand = grids[c].area(px[i, c], py[i, c], radius, 0, lvl)
            elif gate == 3:
                cand = [j for j in grids[c].area(px[i, c], py[i, c], radius, -1, -1) if lvl - 1 <= ck['octave'][j] <= lvl]
            else:
                cand = grids[c].area(px[i, c], py[i, c], radius, lvl - 1, lvl + 1)
            if c == 0 and not cand:
                break
            best, bi = 256, -1
            for j in cand:
                g = j + cam_off[c]
                if occ[g]:
                    continue
                d = hamming(pdesc[i], desc[g])
                if d < best:
                    best, bi = d, j
            if best <= max_dist:
                g = bi + cam_off[c]
                assigned[g] = i; occ[g] = has_obs[i]; nm += 1
                if check:
                    hist[rot_bin(ang[i], ck['angle'][bi])].append(g)
    if check:
        keep = three_maxima_filter(hist)
        for b in range(30):
            if b not in keep:
                for g in hist[b]:
                    assigned[g] = -1; nm -= 1
    # ---- C++ oracle
    L = O.lib()
    cg = (C.c_void_p * 4)()
    for c in range(ncam):
        sub = np.ascontiguousarray(kps[cam_off[c]:cam_off[c + 1]])
        cg[c] = L.orc_grid_create(O.p(sub), sub.size, C.c_float(0), C.c_float(120), C.c_float(0), C.c_float(120))
    ref = np.zeros(N, np.int32)
    rn = L.orc_search_by_projection_last(ncam, O.p(cam_off), O.p(kps), O.p(desc), cg, O.p(sf), O.p(occupied), None, n, O.p(valid), O.p(px), O.p(py),
                                         None, O.p(octv), O.p(ang), O.p(has_obs), O.p(pdesc), C.c_float(th), gate, int(check), max_dist, O.p(ref))
    for c in range(ncam):
        L.orc_grid_destroy(C.c_void_p(cg[c]))
    assert rn == nm and np.array_equal(ref, assigned)
    assert nm > 10


def feature_vector(desc, nbits):
    node = desc[:, 0].astype(np.int32) & ((1 << nbits) - 1)
    order = np.argsort(node, kind='stable')
    nodes, counts = np.unique(node, return_counts=True)
    return nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)


def bow_case(seed, n1, n2, nbits=3):
    rng = np.random.default_rng(seed)
    d2 = synth_descriptors(n2, seed + 1); d2[5:9] = d2[5]
    src = rng.integers(0, n2, n1)
    flips = np.packbits(rng.random((n1, 256)) < 0.04, axis=1, bitorder='little'); flips[:, 0] &= ~np.uint8((1 << nbits) - 1)
    d1 = d2[src] ^ flips
    fresh = rng.random(n1) < 0.3
    d1[fresh] = synth_descriptors(int(fresh.sum()), seed + 2)
    a2 = rng.uniform(0, 360, n2).astype(np.float32)
    a1 = ((a2[src] + np.where(rng.random(n1) < 0.8, 40 + rng.uniform(-4, 4, n1), rng.uniform(0, 360, n1))) % 360).astype(np.float32)
    fv1, fv2 = feature_vector(d1, nbits), feature_vector(d2, nbits)
    keep = [i for i in range(len(fv1[0])) if i % 3 != 1]
    fv1 = (fv1[0][keep].copy(), np.concatenate([[0], np.cumsum(np.diff(fv1[1])[keep])]).astype(np.int32),
           np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in keep]).astype(np.int32))
    return d1, d2, a1, a2, fv1, fv2, (rng.random(n1) < 0.2).astype(np.uint8), (rng.random(n2) < 0.15).astype(np.uint8)


def shared_nodes(fv1, fv2):
    pos2 = {int(v): k for k, v in enumerate(fv2[0])}
    return [(a, pos2[int(v)]) for a, v in enumerate(fv1[0]) if int(v) in pos2]


@pytest.mark.parametrize('seed,nnratio,check', [(11, 0.8, True), (12, 0.6, False), (13, 0.95, True)])
def test_search_by_bow_keyframes_two_restatements(seed, nnratio, check):
    n1, n2 = 260, 240
    d1, d2, a1, a2, fv1, fv2, skip1, skip2 = bow_case(seed, n1, n2)
    match = np.full(n1, -1, np.int32); matched2 = np.zeros(n2, bool); hist = [[] for _ in range(30)]; nm = 0
    for a, b in shared_nodes(fv1, fv2):
        for i1 in fv1[2][fv1[1][a]:fv1[1][a + 1]]:
            if skip1[i1]:
                continue
            b1 = b2 = 256; bi = -1
            for i2 in fv2[2][fv2[1][b]:fv2[1][b + 1]]:
                if matched2[i2] or skip2[i2]:
                    continue
                d = hamming(d1[i1], d2[i2])
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if b1 < 50 and f32(b1) < f32(nnratio) * f32(b2):
                match[i1] = bi; matched2[bi] = True; nm += 1
                if check:
                    hist[rot_bin(a1[i1], a2[bi])].append(i1)
    if check:
        keep = three_maxima_filter(hist)
        for bn in range(30):
            if bn not in keep:
                for i1 in hist[bn]:
                    match[i1] = -1; nm -= 1
    ref = np.zeros(n1, np.int32)
    rn = O.lib().orc_search_by_bow_keyframes(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]), O.p(fv2[2]),
                                             O.p(d1), n1, O.p(d2), n2, O.p(skip1), O.p(skip2), O.p(a1), O.p(a2), C.c_float(f32(nnratio)), int(check),
                                             O.p(ref))
    assert rn == nm and np.array_equal(ref, match) and nm > 30


@pytest.mark.parametrize('seed,ncam,nnratio,check', [(21, 1, 0.7, True), (22, 2, 0.7, True), (23, 4, 0.9, False)])
def test_search_by_bow_keyframe_frame_two_restatements(seed, ncam, nnratio, check):
    n_kf, n_f = 260, 300
    d1, d2, a1, a2, fv1, fv2, skip1, _ = bow_case(seed, n_kf, n_f)
    cam_off = None if ncam == 1 else np.array([0, 110, 200, 260, 290][:ncam + 1], np.int32)
    match = np.full(n_f, -1, np.int32); hist = [[] for _ in range(30)]; nm = 0

    def cam_of(i):
        if ncam == 1:
            return 0
        for c in range(ncam):
            if cam_off[c] <= i < cam_off[c + 1]:
                return c
        return -1
    for a, b in shared_nodes(fv1, fv2):
        for ikf in fv1[2][fv1[1][a]:fv1[1][a + 1]]:
            if skip1[ikf]:
                continue
            b1 = [256] * 4; b2 = [256] * 4; bi = [-1] * 4
            for i_f in fv2[2][fv2[1][b]:fv2[1][b + 1]]:
                if match[i_f] >= 0:
                    continue
                c = cam_of(int(i_f))
                if c < 0:
                    continue
                d = hamming(d1[ikf], d2[i_f])
                if d < b1[c]:
                    b2[c], b1[c], bi[c] = b1[c], d, int(i_f)
                elif d < b2[c]:
                    b2[c] = d
            if b1[0] <= 50:
                for c in range(ncam):
                    ok = (f32(b1[0]) < f32(nnratio) * f32(b2[0])) if c == 0 else (b1[c] <= 50)    # `|| true` for cameras 1-3
                    if not ok:
                        continue
                    match[bi[c]] = ikf; nm += 1
                    if check:
                        hist[rot_bin(a1[ikf], a2[bi[c]])].append(bi[c])
    if check:
        keep = three_maxima_filter(hist)
        for bn in range(30):
            if bn not in keep:
                for i_f in hist[bn]:
                    match[i_f] = -1; nm -= 1
    ref = np.zeros(n_f, np.int32)
    rn = O.lib().orc_search_by_bow_frame(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]), O.p(fv2[2]),
                                         O.p(d1), O.p(d2), n_f, O.p(skip1), O.p(a1), O.p(a2), ncam, None if cam_off is None else O.p(cam_off),
                                         C.c_float(f32(nnratio)), int(check), O.p(ref))
    assert rn == nm and np.array_equal(ref, match) and nm > 20


@pytest.mark.parametrize('seed,window,nnratio,check', [(31, 6, 0.9, True), (32, 15, 0.8, False), (33, 30, 1.0, True)])
def test_search_for_initialization_two_restatements(seed, window, nnratio, check):
    rng = np.random.default_rng(seed)
    n2, n1 = 150, 200
    k2 = random_kps(rng, n2); k2['octave'] = np.where(rng.random(n2) < 0.7, 0, k2['octave'])
    pool = synth_descriptors(5, seed)
    d2 = pooled_desc(rng, n2, pool)
    src = rng.integers(0, n2, n1)
    k1 = k2[src].copy()
    k1['x'] += rng.uniform(-4, 4, n1).astype(np.float32); k1['y'] += rng.uniform(-4, 4, n1).astype(np.float32)
    k1['angle'] = ((k2['angle'][src] + np.where(rng.random(n1) < 0.8, 15.0, rng.uniform(0, 360, n1))) % 360).astype(np.float32)
    d1 = pooled_desc(rng, n1, pool)
    prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1).astype(np.float32)); prev_ref = prev.copy()
    grid = PyGrid(k2, (0, 120, 0, 120))
    m12 = np.full(n1, -1, np.int32); m21 = np.full(n2, -1, np.int32); mdist = np.full(n2, 2 ** 31 - 1, np.int64)
    hist = [[] for _ in range(30)]; nm = 0
    for i1 in range(n1):
        l1 = int(k1['octave'][i1])
        if l1 > 0:
            continue
        cand = grid.area(prev[i1, 0], prev[i1, 1], f32(window), l1, l1)
        if not cand:
            continue
        b1 = b2 = 2 ** 31 - 1; bi = -1
        for i2 in cand:
            d = hamming(d1[i1], d2[i2])
            if mdist[i2] <= d:
                continue
            if d < b1:
                b2, b1, bi = b1, d, i2
            elif d < b2:
                b2 = d
        if b1 <= 50 and b1 < float(f32(b2) * f32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1; nm -= 1
            m12[i1] = bi; m21[bi] = i1; mdist[bi] = b1; nm += 1
            if check:
                hist[rot_bin(k1['angle'][i1], k2['angle'][bi])].append(i1)
    if check:
        keep = three_maxima_filter(hist)
        for bn in range(30):
            if bn not in keep:
                for i1 in hist[bn]:
                    if m12[i1] >= 0:
                        m12[i1] = -1; nm -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = (k2['x'][m12[i1]], k2['y'][m12[i1]])
    L = O.lib()
    g2 = L.orc_grid_create(O.p(k2), n2, C.c_float(0), C.c_float(120), C.c_float(0), C.c_float(120))
    ref = np.zeros(n1, np.int32)
    rn = L.orc_search_for_initialization(O.p(k1), O.p(d1), n1, C.c_void_p(g2), O.p(k2), O.p(d2), n2, O.p(prev_ref), window,
                                         C.c_float(f32(nnratio)), int(check), O.p(ref))
    L.orc_grid_destroy(C.c_void_p(g2))
    assert rn == nm and np.array_equal(ref, m12) and np.array_equal(prev, prev_ref) and nm > 5


@pytest.mark.parametrize('seed,ncam,th,nnratio,far', [(41, 1, 1.0, 0.8, False), (42, 2, 3.0, 0.8, True), (43, 4, 3.0, 0.6, False), (44, 2, 6.0, 1.0, False)])
def test_search_by_projection_map_points_two_restatements(seed, ncam, th, nnratio, far):
    """SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints) src/ORBmatcher.cc:23-340: radius by viewing cosine
    (th only on camera 0), best / second with levels, the ratio-test `continue` that skips the remaining cameras, stereo mirrors."""
    rng = np.random.default_rng(seed)
    ns = [int(rng.integers(40, 90)) for _ in range(ncam)]
    cam_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    N = int(cam_off[-1])
    kps = random_kps(rng, N); pool = synth_descriptors(5, seed)
    desc = pooled_desc(rng, N, pool)
    sf = np.array([f32(1.2) ** l for l in range(8)], np.float32)
    occupied = (rng.random(N) < 0.1).astype(np.uint8)
    n = 160
    tgt = rng.integers(0, N, n)
    px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
    for c in range(ncam):
        kk = np.where((tgt >= cam_off[c]) & (tgt < cam_off[c + 1]), tgt, rng.integers(cam_off[c], cam_off[c + 1], n))
        px[:, c] = kps['x'][kk] + rng.uniform(-2, 2, n); py[:, c] = kps['y'][kk] + rng.uniform(-2, 2, n)
    in_view = (rng.random((n, 4)) < 0.8).astype(np.uint8); in_view[:, ncam:] = 0
    level = np.repeat(np.clip(kps['octave'][tgt] + rng.integers(0, 2, n), 0, 3)[:, None], 4, 1).astype(np.int32)
    level[rng.random((n, 4)) < 0.05] = -1
    level[:, 0] = np.maximum(level[:, 0], 0)                                   # camera 0 assumes a valid level
    vc = rng.choice([0.9990, 0.99], (n, 4)).astype(np.float32)
    depth = rng.uniform(1, 80, n).astype(np.float32); bad = (rng.random(n) < 0.05).astype(np.uint8)
    has_obs = (rng.random(n) < 0.8).astype(np.uint8); pdesc = pooled_desc(rng, n, pool)
    l2r = r2l = None
    if ncam >= 2:
        nl, nr = int(cam_off[1]), int(cam_off[2] - cam_off[1])
        l2r = np.full(nl, -1, np.int32); r2l = np.full(nr, -1, np.int32)
        k = min(nl, nr) // 2
        for a, b in zip(rng.permutation(nl)[:k], rng.permutation(nr)[:k]):
            l2r[a] = b; r2l[b] = a
    # ---- Python restatement
    grids = [PyGrid(kps[cam_off[c]:cam_off[c + 1]], (0, 120, 0, 120)) for c in range(ncam)]
    occ = occupied.copy(); assigned = np.full(N, -1, np.int32); nm = 0
    for m in range(n):
        if not in_view[m].any() or (far and depth[m] > f32(40.0)) or bad[m]:
            continue
        for c in range(4):
            if not in_view[m, c] or c >= ncam:
                continue
            lvl = int(level[m, c])
            if c >= 1 and lvl == -1:
                continue
            r = f32(2.5) if float(vc[m, c]) > 0.998 else f32(4.0)
            if c == 0 and th != 1.0:
                r = f32(r * f32(th))
            ck = kps[cam_off[c]:cam_off[c + 1]]
            cand = grids[c].area(px[m, c], py[m, c], f32(r * sf[lvl]), lvl - 1, lvl)
            if not cand:
                continue
            b1 = b2 = 256; l1 = l2 = -1; bi = -1
            for j in cand:
                if occ[j + cam_off[c]]:
                    continue
                d = hamming(pdesc[m], desc[j + cam_off[c]])
                if d < b1:
                    b2, b1, l2, l1, bi = b1, d, l1, int(ck['octave'][j]), j
                elif d < b2:
                    l2, b2 = int(ck['octave'][j]), d
            if b1 <= 100:
                if l1 == l2 and b1 > float(f32(nnratio) * f32(b2)):
                    break                                                      # `continue` of the map-point loop
                g = bi + cam_off[c]
                assigned[g] = m; occ[g] = has_obs[m]; nm += 1
                if c == 0 and ncam >= 2 and l2r[bi] != -1:
                    gr = l2r[bi] + cam_off[1]
                    assigned[gr] = m; occ[gr] = has_obs[m]; nm += 1
                if c == 1 and r2l[bi] != -1:
                    gl = r2l[bi]
                    assigned[gl] = m; occ[gl] = has_obs[m]; nm += 1
    # ---- C++ oracle
    L = O.lib()
    cg = (C.c_void_p * 4)()
    for c in range(ncam):
        sub = np.ascontiguousarray(kps[cam_off[c]:cam_off[c + 1]])
        cg[c] = L.orc_grid_create(O.p(sub), sub.size, C.c_float(0), C.c_float(120), C.c_float(0), C.c_float(120))
    ref = np.zeros(N, np.int32)
    L.orc_search_by_projection.restype = C.c_int
    rn = L.orc_search_by_projection(ncam, O.p(cam_off), O.p(kps), O.p(desc), cg, O.p(sf), O.p(occupied), None if l2r is None else O.p(l2r),
                                    None if r2l is None else O.p(r2l), None, n, O.p(in_view), O.p(px), O.p(py), O.p(level), O.p(vc), None,
                                    O.p(depth), O.p(bad), O.p(has_obs), O.p(pdesc), C.c_float(th), int(far), C.c_float(40.0),
                                    C.c_float(f32(nnratio)), O.p(ref))
    for c in range(ncam):
        L.orc_grid_destroy(C.c_void_p(cg[c]))
    assert rn == nm and np.array_equal(ref, assigned)
    assert nm > 10


@pytest.mark.parametrize('mb,mbf', [(0.11, 47.9), (0.05, 20.0)])
def test_compute_stereo_matches_two_restatements(mb, mbf):
    """Frame::ComputeStereoMatches src/Frame.cc:1057-1219 restated in Python (float32 step by step) against the C++ oracle, on an
    oracle-extracted synthetic stereo pair."""
    from openmavis_b200.synth import synth_image
    w, h = 376, 240
    left = synth_image(w, h, 2101)
    xs = np.arange(w)
    right = left[:, np.clip(xs + (9 + 6 * np.sin(xs / 60.0)).astype(np.int64), 0, w - 1)].copy()
    EL = O.OracleExtractor(600, 1.2, 8, 20, 7); ER = O.OracleExtractor(600, 1.2, 8, 20, 7)
    _, kl, dl = EL(left, (0, 0)); _, kr, dr = ER(right, (0, 0))
    tab = EL.tables()
    sf, inv_sf = tab['sf'], tab['inv_sf']
    pyrL = [EL.level(l) for l in range(8)]; pyrR = [ER.level(l) for l in range(8)]
    N = kl.size
    u_right = np.full(N, -1, np.float32); depth = np.full(N, -1, np.float32)
    rows = [[] for _ in range(h)]
    for iR in range(kr.size):
        r = f32(f32(2.0) * sf[kr['octave'][iR]])
        maxr = int(math.ceil(float(f32(kr['y'][iR] + r)))); minr = int(math.floor(float(f32(kr['y'][iR] - r))))
        for yi in range(minr, maxr + 1):
            rows[yi].append(iR)
    maxD = f32(f32(mbf) / f32(mb))
    dist_idx = []

    def c_round(v):
        v = float(v)
        return f32(math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5))
    with np.errstate(all='ignore'):
        for iL in range(N):
            lvl = int(kl['octave'][iL]); vL = kl['y'][iL]; uL = kl['x'][iL]
            cands = rows[int(vL)]
            if not cands:
                continue
            minU = f32(uL - maxD); maxU = uL
            if maxU < 0:
                continue
            best, best_r = 100, 0
            for iR in cands:
                if kr['octave'][iR] < lvl - 1 or kr['octave'][iR] > lvl + 1:
                    continue
                if minU <= kr['x'][iR] <= maxU:
                    d = hamming(dl[iL], dr[iR])
                    if d < best:
                        best, best_r = d, iR
            if best >= (100 + 50) // 2:
                continue
            uR0 = kr['x'][best_r]; s = inv_sf[lvl]
            su = c_round(f32(uL * s)); sv = c_round(f32(vL * s)); sr0 = c_round(f32(uR0 * s))
            wdw, L = 5, 5
            IL_ = pyrL[lvl][19 + int(sv) - wdw:19 + int(sv) + wdw + 1, 19 + int(su) - wdw:19 + int(su) + wdw + 1].astype(np.int64)
            cols = pyrR[lvl].shape[1] - 38
            iniu = f32(sr0 + f32(L) - f32(wdw)); endu = f32(sr0 + f32(L) + f32(wdw) + f32(1))
            if iniu < 0 or endu >= cols:
                continue
            bestd, bestinc = 2 ** 31 - 1, 0
            vd = np.zeros(2 * L + 1, np.float32)
            for inc in range(-L, L + 1):
                x0 = 19 + int(sr0) + inc - wdw
                IR_ = pyrR[lvl][19 + int(sv) - wdw:19 + int(sv) + wdw + 1, x0:x0 + 2 * wdw + 1].astype(np.int64)
                dist = f32(np.abs(IL_ - IR_).sum())
                if dist < f32(bestd):
                    bestd, bestinc = int(dist), inc
                vd[L + inc] = dist
            if bestinc in (-L, L):
                continue
            d1, d2, d3 = vd[L + bestinc - 1], vd[L + bestinc], vd[L + bestinc + 1]
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
            if delta < -1 or delta > 1:
                continue
            best_u = f32(sf[lvl] * f32(f32(sr0 + f32(bestinc)) + delta))
            disp = f32(uL - best_u)
            if disp >= 0 and disp < maxD:
                if disp <= 0:
                    disp = f32(0.01); best_u = f32(float(uL) - 0.01)
                depth[iL] = f32(f32(mbf) / disp); u_right[iL] = best_u
                dist_idx.append((bestd, iL))
    dist_idx.sort()
    median = f32(dist_idx[len(dist_idx) // 2][0])
    th = f32(f32(f32(1.5) * f32(1.4)) * median)
    for d, i in reversed(dist_idx):
        if f32(d) < th:
            break
        u_right[i] = -1; depth[i] = -1
    ru = np.zeros(N, np.float32); rd = np.zeros(N, np.float32)
    L_ = O.lib()
    L_.orc_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = L_.orc_compute_stereo_matches(EL.h, ER.h, mb, mbf, O.p(ru), O.p(rd))
    assert n == int((depth > 0).sum()) and n > 60
    assert np.array_equal(ru.view(np.uint32), u_right.view(np.uint32)) and np.array_equal(rd.view(np.uint32), depth.view(np.uint32))


def test_distinctive_descriptors_and_vocabulary_two_restatements():
    """MapPoint::ComputeDistinctiveDescriptors src/MapPoint.cc:405-483 and the DBoW2 descent (TemplatedVocabulary.h:1213-1259)."""
    rng = np.random.default_rng(7)
    sizes = rng.integers(0, 14, 120).astype(np.int32); sizes[::11] = 1
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    pool = synth_descriptors(120, 3)
    desc = np.repeat(pool, sizes, axis=0) ^ np.packbits(rng.random((int(ptr[-1]), 256)) < 0.1, axis=1, bitorder='little')
    best = np.zeros(120, np.int32)
    O.lib().orc_distinctive_descriptors(O.p(desc), O.p(ptr), 120, O.p(best))
    for g in range(120):
        n = int(sizes[g])
        if n == 0:
            assert best[g] == -1
            continue
        rows = desc[ptr[g]:ptr[g + 1]]
        D = np.array([[hamming(a, b) for b in rows] for a in rows])
        med = [sorted(D[i])[int(0.5 * (n - 1))] for i in range(n)]
        assert best[g] == int(np.argmin(med))                                  # first minimum
    # vocabulary: k-ary tree with random node descriptors; the descent takes the FIRST child with the least distance
    k, Lv = 5, 3
    # (breadth-first numbering)
    nodes = [dict(children=[], desc=np.zeros(32, np.uint8), word=-1)]
    frontier = [0]
    for lev in range(Lv):
        nxt = []
        for nid in frontier:
            for _ in range(k):
                nodes.append(dict(children=[], desc=synth_descriptors(1, len(nodes) + 100)[0], word=-1))
                nodes[nid]['children'].append(len(nodes) - 1); nxt.append(len(nodes) - 1)
        frontier = nxt
    for wid, nid in enumerate(frontier):
        nodes[nid]['word'] = wid
    nodes[3]['desc'] = nodes[2]['desc'].copy()                                  # a tie between siblings
    child_ptr = np.zeros(len(nodes) + 1, np.int32); child = []
    for i, nd in enumerate(nodes):
        child += nd['children']; child_ptr[i + 1] = len(child)
    child = np.array(child, np.int32); ndesc = np.stack([nd['desc'] for nd in nodes]); word = np.array([nd['word'] for nd in nodes], np.int32)
    feats = synth_descriptors(300, 9); feats[:20] = ndesc[rng.integers(1, len(nodes), 20)]
    ow = np.zeros(300, np.int32); on = np.zeros(300, np.int32)
    O.lib().orc_vocab_transform(O.p(child_ptr), O.p(child), O.p(ndesc), O.p(word), Lv, O.p(feats), 300, 2, O.p(ow), O.p(on))
    for i in range(300):
        nid, lev, node_at = 0, 0, 0
        while nodes[nid]['children']:
            ch = nodes[nid]['children']
            ds = [hamming(feats[i], nodes[c]['desc']) for c in ch]
            nid = ch[int(np.argmin(ds))]; lev += 1
            if lev == Lv - 2:
                node_at = nid
        assert ow[i] == nodes[nid]['word'] and on[i] == node_at


def test_fuse_search_loop_sim3_and_triangulation_two_restatements():
    """Fuse's search loop (src/ORBmatcher.cc:1549-1601 with the 5.99 / 7.8 reprojection gates), SearchBySim3 (:1771-1983) and the
    Hamming core of SearchForTriangulation (:1238-1290, bCoarse: LAST minimum within TH_LOW)."""
    rng = np.random.default_rng(51)
    L = O.lib()
    n, nq = 140, 160
    kps = random_kps(rng, n); pool = synth_descriptors(5, 51); desc = pooled_desc(rng, n, pool)
    sf = np.array([f32(1.2) ** l for l in range(8)], np.float32)
    inv_s2 = (f32(1.0) / (sf * sf)).astype(np.float32)
    tgt = rng.integers(0, n, nq)
    qx = (kps['x'][tgt] + rng.uniform(-3, 3, nq)).astype(np.float32); qy = (kps['y'][tgt] + rng.uniform(-3, 3, nq)).astype(np.float32)
    lvl = np.clip(kps['octave'][tgt] + rng.integers(0, 2, nq), 0, 3).astype(np.int32); mn = (lvl - 1).astype(np.int32)
    radius = (f32(3.0) * sf[lvl]).astype(np.float32)
    qd = pooled_desc(rng, nq, pool); valid = (rng.random(nq) < 0.9).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, kps['x'] - rng.uniform(1, 20, n), -1).astype(np.float32)
    q_ur = (qx - rng.uniform(0, 22, nq)).astype(np.float32)
    grid = PyGrid(kps, (0, 120, 0, 120))
    g = L.orc_grid_create(O.p(kps), n, C.c_float(0), C.c_float(120), C.c_float(0), C.c_float(120))
    for gates in (True, False):
        bi = np.full(nq, -1, np.int32); bd = np.full(nq, 256, np.int32)
        for q in range(nq):
            if not valid[q]:
                continue
            for j in grid.area(qx[q], qy[q], radius[q], -1, -1):
                o = int(kps['octave'][j])
                if o < mn[q] or o > lvl[q]:
                    continue
                if gates:
                    ex = f32(qx[q] - kps['x'][j]); ey = f32(qy[q] - kps['y'][j])
                    if u_right[j] >= 0:
                        er = f32(q_ur[q] - u_right[j])
                        e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                        if float(f32(e2 * inv_s2[o])) > 7.8:
                            continue
                    else:
                        e2 = f32(f32(ex * ex) + f32(ey * ey))
                        if float(f32(e2 * inv_s2[o])) > 5.99:
                            continue
                d = hamming(qd[q], desc[j])
                if d < bd[q]:
                    bd[q], bi[q] = d, j
        ri = np.zeros(nq, np.int32); rd = np.zeros(nq, np.int32)
        L.orc_window_best(C.c_void_p(g), O.p(kps), O.p(desc), None, O.p(u_right) if gates else None, O.p(inv_s2) if gates else None, nq,
                          O.p(valid), O.p(qx), O.p(qy), O.p(q_ur) if gates else None, O.p(radius), O.p(mn), O.p(lvl), O.p(qd), O.p(ri), O.p(rd))
        assert np.array_equal(ri, bi) and np.array_equal(rd, bd)
    # ---- SearchBySim3 on (kps, desc) x (kps2, desc2)
    n2 = 130
    k2 = random_kps(rng, n2); d2 = pooled_desc(rng, n2, pool); grid2 = PyGrid(k2, (0, 120, 0, 120))
    g2 = L.orc_grid_create(O.p(k2), n2, C.c_float(0), C.c_float(120), C.c_float(0), C.c_float(120))

    def side(n_self, k_other, target):
        u = (k_other['x'][target] + rng.uniform(-2, 2, n_self)).astype(np.float32); v = (k_other['y'][target] + rng.uniform(-2, 2, n_self)).astype(np.float32)
        lv = np.clip(k_other['octave'][target] + rng.integers(0, 2, n_self), 0, 3).astype(np.int32)
        return dict(valid=(rng.random(n_self) < 0.85).astype(np.uint8), u=u, v=v, radius=(f32(7.5) * sf[lv]).astype(np.float32), level=lv,
                    desc=pooled_desc(rng, n_self, pool))
    s1 = side(n, k2, rng.integers(0, n2, n)); s2 = side(n2, kps, rng.integers(0, n, n2))

    def one_way(s, grid_o, k_o, d_o):
        out = np.full(s['u'].size, -1, np.int64)
        for i in range(s['u'].size):
            if not s['valid'][i]:
                continue
            best, bj = 2 ** 31 - 1, -1
            for j in grid_o.area(s['u'][i], s['v'][i], s['radius'][i], -1, -1):
                if k_o['octave'][j] < s['level'][i] - 1 or k_o['octave'][j] > s['level'][i]:
                    continue
                d = hamming(s['desc'][i], d_o[j])
                if d < best:
                    best, bj = d, j
            if best <= 100:
                out[i] = bj
        return out
    m1 = one_way(s1, grid2, k2, d2); m2 = one_way(s2, grid, kps, desc)
    m12 = np.array([m1[i] if m1[i] >= 0 and m2[m1[i]] == i else -1 for i in range(n)], np.int32)
    ref = np.zeros(n, np.int32)
    rn = L.orc_search_by_sim3(C.c_void_p(g), O.p(kps), O.p(desc), n, C.c_void_p(g2), O.p(k2), O.p(d2), n2, O.p(s1['valid']), O.p(s1['u']), O.p(s1['v']),
                              O.p(s1['radius']), O.p(s1['level']), O.p(s1['desc']), O.p(s2['valid']), O.p(s2['u']), O.p(s2['v']), O.p(s2['radius']),
                              O.p(s2['level']), O.p(s2['desc']), O.p(ref))
    L.orc_grid_destroy(C.c_void_p(g)); L.orc_grid_destroy(C.c_void_p(g2))
    assert rn == int((m12 >= 0).sum()) and np.array_equal(ref, m12)
    # ---- SearchForTriangulation core
    d1, dd2, _, _, fv1, fv2, skip1, skip2 = bow_case(61, 220, 200)
    match = np.full(220, -1, np.int32)
    for a, b in shared_nodes(fv1, fv2):
        for i1 in fv1[2][fv1[1][a]:fv1[1][a + 1]]:
            if skip1[i1]:
                continue
            best, bj = 50, -1
            for i2 in fv2[2][fv2[1][b]:fv2[1][b + 1]]:
                if skip2[i2]:
                    continue
                d = hamming(d1[i1], dd2[i2])
                if d > 50 or d > best:
                    continue
                bj, best = int(i2), d
            match[i1] = bj
    ref = np.zeros(220, np.int32)
    rn = L.orc_search_for_triangulation(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]), O.p(fv2[2]),
                                        O.p(d1), 220, O.p(dd2), O.p(skip1), O.p(skip2), None, None, O.p(ref))
    assert rn == int((match >= 0).sum()) and np.array_equal(ref, match) and rn > 20
