"""GPU parity tests of the Hamming-matching kernels vs the CPU oracle (bit-exact indices and distances).
Reference: src/ORBmatcher.cc (DescriptorDistance :2577, SearchByProjection :23-340, SearchForTriangulation :1131-1456),
cv::BFMatcher::knnMatch at src/Frame.cc:1427,1483, Frame grid src/Frame.cc:541-582, 890-967."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import pytest

import oracle_lib as O
import openmavis_b200 as ob
from openmavis_b200.synth import synth_descriptors, synth_image, synth_query_from_train

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def M():
    if ob.device_count() == 0:
        pytest.fail('no CUDA device: -m gpu tests must run on the GPU box')
    return ob.ORBmatcher(0.8)


def test_descriptor_distance(M):
    a = synth_descriptors(1000, 1); b = synth_descriptors(1000, 2)
    b[:10] = a[:10]; b[10:20] = ~a[10:20]
    d = ob.ORBmatcher.DescriptorDistance(a, b)
    ref = np.array([O.hamming(a[i], b[i]) for i in range(1000)])
    assert np.array_equal(d, ref) and d[0] == 0 and d[10] == 256
    assert ob.ORBmatcher.DescriptorDistance(a[5], b[500]) == ref_single(a[5], b[500])


def ref_single(a, b):
    return O.hamming(a, b)


@pytest.mark.parametrize('nq,nt', [(1, 1), (1, 2), (7, 3), (300, 1), (1500, 1500), (1537, 2049), (40, 9000), (5000, 257)])
def test_knn2_matches_oracle(M, nq, nt):
    t = synth_descriptors(nt, 100 + nt)
    t[: nt // 3, 6:] = 0                       # distance ties between different train rows
    q = synth_query_from_train(np.resize(t, (max(nq, nt), 32))[:max(nq, nt)], 7)[:nq]
    idx, dist = M.knnMatch2(q, t)
    ridx, rdist = O.knn2(q, t)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    assert np.array_equal(ob.ORBmatcher.ratio_test(dist, 0.8), O.ratio_test(rdist, 0.8))


def test_knn2_empty_inputs(M):
    t = synth_descriptors(10, 1)
    idx, dist = M.knnMatch2(np.zeros((0, 32), np.uint8), t)
    assert idx.shape == (0, 2)
    idx, dist = M.knnMatch2(t, np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (dist == -1).all()


def test_knn2_100k_property(M):
    """Config 5 size (100k x 100k): too big for the scalar oracle -> size-independent properties: every planted
    neighbour is recovered with its exact distance, results are (dist asc, idx asc), and a random sample of rows
    equals the oracle's full scan."""
    n = 100000
    train = synth_descriptors(n, 5)
    rng = np.random.default_rng(6)
    src = rng.integers(0, n, size=n)
    flips = np.packbits(rng.random((n, 256)) < 0.04, axis=1, bitorder='little')
    q = train[src] ^ flips
    idx, dist = M.knnMatch2(q, train)
    planted = np.unpackbits(flips, axis=1).sum(axis=1)
    d_to_src = np.unpackbits(q ^ train[src], axis=1).sum(axis=1)
    assert np.array_equal(planted, d_to_src)
    assert (dist[:, 0] <= planted).all()
    same = idx[:, 0] == src
    assert same.mean() > 0.999 and np.array_equal(dist[same, 0], planted[same])
    assert (dist[:, 0] <= dist[:, 1]).all()
    tie = dist[:, 0] == dist[:, 1]
    assert (idx[tie, 0] < idx[tie, 1]).all()
    got_d0 = np.unpackbits(q ^ train[idx[:, 0]], axis=1).sum(axis=1)
    got_d1 = np.unpackbits(q ^ train[idx[:, 1]], axis=1).sum(axis=1)
    assert np.array_equal(got_d0, dist[:, 0]) and np.array_equal(got_d1, dist[:, 1])
    rows = rng.integers(0, n, size=48)
    ridx, rdist = O.knn2(q[rows], train)
    assert np.array_equal(idx[rows], ridx) and np.array_equal(dist[rows], rdist)


def test_knn2_100k_config5_generator(M):
    """The 100k x 100k workload exactly as SURVEY.md §8(d) config 5 defines it and bench.py times it (train = uniform random rows of
    seed 5; query i = a train row with Binomial(256, 0.08) bit flips for 80 % of the rows, fresh random rows for 20 %: real nearest
    neighbours AND distance ties): size-independent properties + a sampled full scan by the oracle."""
    from openmavis_b200.synth import synth_query_from_train
    n = 100000
    train = synth_descriptors(n, 5)
    q = synth_query_from_train(train, 6)
    idx, dist = M.knnMatch2(q, train)
    assert idx.shape == (n, 2) and (idx >= 0).all() and (idx < n).all()
    got_d0 = np.unpackbits(q ^ train[idx[:, 0]], axis=1).sum(axis=1)
    got_d1 = np.unpackbits(q ^ train[idx[:, 1]], axis=1).sum(axis=1)
    assert np.array_equal(got_d0, dist[:, 0]) and np.array_equal(got_d1, dist[:, 1])          # reported distances are the true ones
    assert (dist[:, 0] <= dist[:, 1]).all() and (idx[:, 0] != idx[:, 1]).all()
    tie = dist[:, 0] == dist[:, 1]
    assert (idx[tie, 0] < idx[tie, 1]).all()                                                  # (distance asc, index asc)
    near = dist[:, 0] < 60                                                                     # the planted 80 %: Binomial(256, 0.08) stays below 60
    assert 0.78 < near.mean() < 0.82
    assert (dist[~near, 0] >= 60).all() and dist[~near, 0].min() >= 70                         # fresh rows: nearest of 100k random rows
    rows = np.concatenate([np.random.default_rng(7).integers(0, n, size=40), np.flatnonzero(tie)[:8]])
    ridx, rdist = O.knn2(q[rows], train)
    assert np.array_equal(idx[rows], ridx) and np.array_equal(dist[rows], rdist)


def _oracle_csr(q, t, row_ptr, cand):
    nq = q.shape[0]
    out = dict(best_idx=np.full(nq, -1, np.int32), best_dist=np.full(nq, 256, np.int32), second_dist=np.full(nq, 256, np.int32),
               best_pos=np.full(nq, -1, np.int32), second_pos=np.full(nq, -1, np.int32))
    dists = np.zeros(len(cand), np.int32)
    for i in range(nq):
        bd, bd2, bi, bp, bp2 = 256, 256, -1, -1, -1
        for k in range(row_ptr[i], row_ptr[i + 1]):
            d = O.hamming(q[i], t[cand[k]]); dists[k] = d
            if d < bd:
                bd2, bp2, bd, bi, bp = bd, bp, d, cand[k], k - row_ptr[i]
            elif d < bd2:
                bd2, bp2 = d, k - row_ptr[i]
        out['best_idx'][i], out['best_dist'][i], out['second_dist'][i], out['best_pos'][i], out['second_pos'][i] = bi, bd, bd2, bp, bp2
    return out, dists


def test_match_csr(M):
    rng = np.random.default_rng(3)
    t = synth_descriptors(800, 31); t[:200, 3:] = 0
    q = synth_query_from_train(t, 32)[:500]
    lens = rng.integers(0, 40, size=500); lens[::9] = 0
    row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, 800, size=int(row_ptr[-1])).astype(np.int32)
    got = M.match_csr(q, t, row_ptr, cand)
    ref, rd = _oracle_csr(q, t, row_ptr, cand)
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k
    assert np.array_equal(M.match_csr_dists(q, t, row_ptr, cand), rd)


def _feature_vector(desc, nbits=6):
    """Stand-in for DBoW2 level-4 nodes (vocabulary file absent): node = first nbits of the descriptor."""
    node = desc[:, 0].astype(np.int32) & ((1 << nbits) - 1)
    order = np.argsort(node, kind='stable')
    nodes, counts = np.unique(node, return_counts=True)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return nodes.astype(np.int32), ptr, order.astype(np.int32)


def test_search_for_triangulation(M):
    rng = np.random.default_rng(4)
    d2 = synth_descriptors(1200, 41)
    d2[100:160] = d2[100]                              # exact duplicates -> last-minimum tie rule
    src = rng.integers(0, 1200, 1200)
    flips = np.packbits(rng.random((1200, 256)) < 0.05, axis=1, bitorder='little')
    flips[:, 0] &= 0xC0                                # keep the node bits: d1[i] shares the node of d2[src[i]]
    d1 = d2[src] ^ flips
    d1[::5] = synth_descriptors(240, 43)               # unrelated rows
    fv1, fv2 = _feature_vector(d1), _feature_vector(d2)
    fv1 = (fv1[0][::2].copy(), np.concatenate([[0], np.cumsum(np.diff(fv1[1])[::2])]).astype(np.int32),
           np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in range(0, len(fv1[0]), 2)]).astype(np.int32))  # nodes only in KF2
    skip1 = (rng.random(1200) < 0.3).astype(np.uint8); skip2 = (rng.random(1200) < 0.3).astype(np.uint8)
    nm, pairs, m12 = M.SearchForTriangulation(fv1, fv2, d1, d2, skip1, skip2)
    ref = np.zeros(1200, np.int32)
    L = O.lib()
    rn = L.orc_search_for_triangulation(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]),
                                        O.p(fv2[2]), O.p(d1), 1200, O.p(d2), O.p(skip1), O.p(skip2), None, None, O.p(ref))
    assert rn == nm and np.array_equal(ref, m12) and nm > 150, (rn, nm)
    assert pairs == [(int(i), int(ref[i])) for i in np.nonzero(ref >= 0)[0]]


def _bow_case(seed, n1, n2, nbits):
    """Two feature sets sharing vocabulary nodes: 70 % of set 1 are noisy copies of set-2 rows (node bits kept), several copies of
    the same row compete for it (the already-matched rule), exact duplicates in set 2 tie on distance."""
    rng = np.random.default_rng(seed)
    d2 = synth_descriptors(n2, seed + 1)
    d2[50:58] = d2[50]
    src = rng.integers(0, n2, n1)
    flips = np.packbits(rng.random((n1, 256)) < 0.04, axis=1, bitorder='little')
    flips[:, 0] &= ~np.uint8((1 << nbits) - 1)
    d1 = d2[src] ^ flips
    fresh = rng.random(n1) < 0.3
    d1[fresh] = synth_descriptors(int(fresh.sum()), seed + 2)
    a2 = rng.uniform(0, 360, n2).astype(np.float32)
    a1 = ((a2[src] + np.where(rng.random(n1) < 0.8, 40 + rng.uniform(-4, 4, n1), rng.uniform(0, 360, n1))) % 360).astype(np.float32)
    fv1, fv2 = _feature_vector(d1, nbits), _feature_vector(d2, nbits)
    # drop every third node of list 1 so that the merge walk has to skip on both sides
    keep = [i for i in range(len(fv1[0])) if i % 3 != 1]
    fv1 = (fv1[0][keep].copy(), np.concatenate([[0], np.cumsum(np.diff(fv1[1])[keep])]).astype(np.int32),
           np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in keep]).astype(np.int32))
    skip1 = (rng.random(n1) < 0.2).astype(np.uint8); skip2 = (rng.random(n2) < 0.15).astype(np.uint8)
    return d1, d2, a1, a2, fv1, fv2, skip1, skip2


@pytest.mark.parametrize('ncam,nnratio,check,nbits', [(1, 0.7, True, 5), (2, 0.7, True, 5), (4, 0.75, True, 6), (4, 0.9, False, 4), (1, 0.6, True, 2)])
def test_search_by_bow_keyframe_frame(ncam, nnratio, check, nbits):
    """src/ORBmatcher.cc:349-666: per-camera best / second, ratio only on camera 0, already-matched rule, rotation filter."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    n_kf, n_f = 2500, 3000
    d1, d2, a1, a2, fv1, fv2, skip1, _ = _bow_case(60 + ncam, n_kf, n_f, nbits)
    cam_off = None if ncam == 1 else np.array([0, 1100, 2000, 2600, 2950][:ncam + 1], np.int32)   # features beyond the last block are ignored
    M = ob.ORBmatcher(nnratio, checkOri=check)
    nm, match = M.SearchByBoW(fv1, fv2, d1, d2, skip1, a1, a2, cam_off)
    ref = np.zeros(n_f, np.int32)
    L = O.lib()
    rn = L.orc_search_by_bow_frame(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]), O.p(fv2[2]),
                                   O.p(d1), O.p(d2), n_f, O.p(skip1), O.p(a1), O.p(a2), ncam, None if cam_off is None else O.p(cam_off),
                                   C.c_float(np.float32(nnratio)), int(check), O.p(ref))
    assert rn == nm, (rn, nm)
    assert np.array_equal(ref, match), int((ref != match).sum())
    assert nm > 150
    assert (skip1[match[match >= 0]] == 0).all()


@pytest.mark.parametrize('nnratio,check,nbits', [(0.8, True, 5), (0.6, False, 6), (0.9, True, 1)])
def test_search_by_bow_keyframes(nnratio, check, nbits):
    """src/ORBmatcher.cc:1006-1129: strict < TH_LOW, ratio test, vbMatched2, rotation filter."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    n1, n2 = 2200, 2000
    d1, d2, a1, a2, fv1, fv2, skip1, skip2 = _bow_case(80 + nbits, n1, n2, nbits)
    M = ob.ORBmatcher(nnratio, checkOri=check)
    nm, m12 = M.SearchByBoWKeyFrames(fv1, fv2, d1, d2, skip1, skip2, a1, a2)
    ref = np.zeros(n1, np.int32)
    L = O.lib()
    rn = L.orc_search_by_bow_keyframes(len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]), O.p(fv1[2]), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]), O.p(fv2[2]),
                                       O.p(d1), n1, O.p(d2), n2, O.p(skip1), O.p(skip2), O.p(a1), O.p(a2), C.c_float(np.float32(nnratio)),
                                       int(check), O.p(ref))
    assert rn == nm, (rn, nm)
    assert np.array_equal(ref, m12), int((ref != m12).sum())
    assert nm > 200
    used = m12[m12 >= 0]
    assert np.unique(used).size == used.size and (skip2[used] == 0).all()


def test_search_by_bow_degenerate(M):
    d = synth_descriptors(10, 1)
    fv = _feature_vector(d, 2)
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    nm, match = M.SearchByBoW(empty, fv, np.zeros((0, 32), np.uint8), d)
    assert nm == 0 and (match == -1).all() and match.size == 10
    nm, m12 = M.SearchByBoWKeyFrames(fv, fv, d, d, angle1=np.zeros(10, np.float32), angle2=np.zeros(10, np.float32))
    assert nm == 10 and np.array_equal(m12, np.arange(10))          # identical sets match themselves (distance 0)
    with pytest.raises(ob.OrbxError):
        M.SearchByBoWKeyFrames((fv[0], fv[1], fv[2] + 5), fv, d, d, angle1=np.zeros(10, np.float32), angle2=np.zeros(10, np.float32))


def _extract_frame(w, h, nf, seeds, lap):
    rig = ob.RigExtractor(w, h, len(seeds), nf)
    imgs = [synth_image(w, h, s) for s in seeds]
    return rig, rig.extract(imgs, ob.make_params(len(seeds), nf, 20, 7, lap))


def test_grid_candidates_match_frame_grid(M):
    rig, res = _extract_frame(512, 512, 1500, [4100], (0, 511))
    kps = res[0][1].copy()
    bounds = (0.0, 512.0, 0.0, 512.0)
    rng = np.random.default_rng(5)
    nq = 3000
    qx = rng.uniform(-20, 530, nq).astype(np.float32); qy = rng.uniform(-20, 530, nq).astype(np.float32)
    lv = rng.integers(0, 8, nq).astype(np.int32)
    qr = (rng.choice([2.5, 4.0, 7.5, 15.0], nq) * rig.scale[lv]).astype(np.float32)
    mn = lv - 1; mx = lv.copy(); mx[::11] = -1; mn[::11] = -1     # also the "no level check" form
    row_ptr, cand = M.grid_candidates(kps, bounds, qx, qy, qr, mn, mx)
    g = O.lib().orc_grid_create(O.p(kps), kps.size, C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
    buf = np.zeros(4096, np.int32)
    for i in range(nq):
        n = O.lib().orc_grid_query(C.c_void_p(g), O.p(kps), C.c_float(qx[i]), C.c_float(qy[i]), C.c_float(qr[i]), int(mn[i]), int(mx[i]),
                                   O.p(buf), 4096)
        assert n == row_ptr[i + 1] - row_ptr[i], i
        assert np.array_equal(buf[:n], cand[row_ptr[i]:row_ptr[i + 1]]), i
    O.lib().orc_grid_destroy(C.c_void_p(g))
    assert row_ptr[-1] > 3000


def _sbp_case(ncam, seed, nmp=5000, th=3.0, nnratio=0.8):
    """Config 4 style: map-point descriptors come from the frame's own keypoints (70 %, with a few
    flipped bits) or are random (30 %); predicted position = keypoint position + U(-3,3)."""
    rng = np.random.default_rng(seed)
    w = h = 512
    rig, res = _extract_frame(w, h, 1500, [4200 + seed * 10 + c for c in range(ncam)], (0, 511))
    kps = np.concatenate([r[1] for r in res]); desc = np.concatenate([r[2] for r in res])
    cam_off = np.concatenate([[0], np.cumsum([len(r[1]) for r in res])]).astype(np.int32)
    N = int(cam_off[-1])
    in_view = np.zeros((nmp, 4), np.uint8); px = np.zeros((nmp, 4), np.float32); py = np.zeros((nmp, 4), np.float32)
    lvl = np.full((nmp, 4), -1, np.int32); vc = np.zeros((nmp, 4), np.float32); md = synth_descriptors(nmp, seed + 1)
    for i in range(nmp):
        cams = [c for c in range(ncam) if rng.random() < (0.9 if c == 0 else 0.5)]
        src_cam = cams[0] if cams else 0
        k = int(rng.integers(cam_off[src_cam], cam_off[src_cam + 1]))
        if rng.random() < 0.7:
            flips = np.packbits(rng.random(256) < 0.06, bitorder='little')
            md[i] = desc[k] ^ flips
        for c in cams:
            kk = k if c == src_cam else int(rng.integers(cam_off[c], cam_off[c + 1]))
            in_view[i, c] = 1
            px[i, c] = kps['x'][kk] + rng.uniform(-3, 3); py[i, c] = kps['y'][kk] + rng.uniform(-3, 3)
            lvl[i, c] = min(7, max(0, int(kps['octave'][kk]) + int(rng.integers(0, 2))))
            vc[i, c] = rng.choice([0.9990, 0.99])
    occupied = (rng.random(N) < 0.05).astype(np.uint8)
    has_obs = (rng.random(nmp) < 0.9).astype(np.uint8)
    bad = (rng.random(nmp) < 0.02).astype(np.uint8)
    depth = rng.uniform(1, 80, nmp).astype(np.float32)
    l2r = r2l = None
    if ncam >= 2:
        nl, nr = int(cam_off[1]), int(cam_off[2] - cam_off[1])
        l2r = np.full(nl, -1, np.int32); r2l = np.full(nr, -1, np.int32)
        for a, b in zip(rng.permutation(nl)[:300], rng.permutation(nr)[:300]):
            l2r[a] = b; r2l[b] = a
    frame = dict(cam_off=cam_off, kps=kps, desc=desc, bounds=np.tile(np.array([0, w, 0, h], np.float32), (ncam, 1)),
                 scale_factors=rig.scale, occupied=occupied, l2r=l2r, r2l=r2l)
    points = dict(in_view=in_view, proj_x=px, proj_y=py, level=lvl, view_cos=vc, desc=md, depth=depth, bad=bad, has_obs=has_obs)
    return frame, points, th, nnratio


def _oracle_sbp(frame, points, th, nnratio, far, thfar):
    L = O.lib()
    cam_off = frame['cam_off']; ncam = len(cam_off) - 1; N = int(cam_off[-1])
    kps = np.ascontiguousarray(frame['kps']); desc = np.ascontiguousarray(frame['desc'])
    grids = (C.c_void_p * 4)()
    for c in range(ncam):
        b = frame['bounds'][c]
        sub = np.ascontiguousarray(kps[cam_off[c]:cam_off[c + 1]])
        grids[c] = L.orc_grid_create(O.p(sub), sub.size, C.c_float(b[0]), C.c_float(b[1]), C.c_float(b[2]), C.c_float(b[3]))
    assigned = np.zeros(N, np.int32)
    L.orc_search_by_projection.restype = C.c_int
    n = L.orc_search_by_projection(ncam, O.p(cam_off), O.p(kps), O.p(desc), grids, O.p(frame['scale_factors']), O.p(frame['occupied']),
                                   None if frame['l2r'] is None else O.p(frame['l2r']), None if frame['r2l'] is None else O.p(frame['r2l']),
                                   None if frame.get('u_right') is None else O.p(frame['u_right']), points['in_view'].shape[0],
                                   O.p(points['in_view']), O.p(points['proj_x']), O.p(points['proj_y']), O.p(points['level']), O.p(points['view_cos']),
                                   None if points.get('proj_xr') is None else O.p(points['proj_xr']), O.p(points['depth']), O.p(points['bad']),
                                   O.p(points['has_obs']), O.p(points['desc']), C.c_float(th), int(far), C.c_float(thfar),
                                   C.c_float(nnratio), O.p(assigned))
    for c in range(ncam):
        L.orc_grid_destroy(C.c_void_p(grids[c]))
    return n, assigned


@pytest.mark.parametrize('ncam,th,far', [(1, 1.0, False), (1, 3.0, True), (2, 3.0, False), (2, 6.0, True), (4, 3.0, False)])
def test_search_by_projection(ncam, th, far):
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, points, _, nnratio = _sbp_case(ncam, seed=ncam * 3 + int(th))
    M = ob.ORBmatcher(nnratio)
    nm, assigned = M.SearchByProjection(frame, points, th, far, 40.0)
    rn, rassigned = _oracle_sbp(frame, points, th, np.float32(nnratio), far, 40.0)
    assert rn == nm, (rn, nm)
    assert np.array_equal(assigned, rassigned), int((assigned != rassigned).sum())
    assert nm > 500


def _sbp_last_case(ncam, seed, n=4000):
    """Current frame = keypoints extracted from `ncam` 512x512 synthetic images; last-frame keypoints with map points observe
    current keypoints (+-3 px, octave +-1, angle = current + 20 deg +- 3 for 85 %, random otherwise)."""
    rng = np.random.default_rng(seed)
    w = h = 512
    rig, res = _extract_frame(w, h, 1500, [4700 + seed * 10 + c for c in range(ncam)], (0, 511))
    kps = np.concatenate([r[1] for r in res]); desc = np.concatenate([r[2] for r in res])
    cam_off = np.concatenate([[0], np.cumsum([len(r[1]) for r in res])]).astype(np.int32)
    N = int(cam_off[-1])
    valid = (rng.random(n) < 0.93).astype(np.uint8)
    px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
    octv = np.zeros(n, np.int32); ang = np.zeros(n, np.float32); pd = synth_descriptors(n, seed + 1)
    for i in range(n):
        k = int(rng.integers(0, N))
        cam = int(np.searchsorted(cam_off, k, side='right') - 1)
        if rng.random() < 0.75:
            pd[i] = desc[k] ^ np.packbits(rng.random(256) < 0.06, bitorder='little')
        for c in range(ncam):
            if c == cam or (c >= 2 and cam == 0):        # the side cameras are searched around the LEFT projection (D6)
                px[i, c] = kps['x'][k] + rng.uniform(-3, 3); py[i, c] = kps['y'][k] + rng.uniform(-3, 3)
            elif c == 0 and rng.random() < 0.5:
                px[i, c] = rng.uniform(-40, 560); py[i, c] = rng.uniform(-40, 560)     # often an empty camera-0 window
            else:
                kk = int(rng.integers(cam_off[c], cam_off[c + 1]))
                px[i, c] = kps['x'][kk] + rng.uniform(-3, 3); py[i, c] = kps['y'][kk] + rng.uniform(-3, 3)
        octv[i] = min(7, max(0, int(kps['octave'][k]) + int(rng.integers(-1, 2))))
        ang[i] = (kps['angle'][k] + (20 + rng.uniform(-3, 3) if rng.random() < 0.85 else rng.uniform(0, 360))) % 360.0
    frame = dict(cam_off=cam_off, kps=kps, desc=desc, bounds=np.tile(np.array([0, w, 0, h], np.float32), (ncam, 1)),
                 scale_factors=rig.scale, occupied=(rng.random(N) < 0.05).astype(np.uint8))
    last = dict(valid=valid, proj_x=px, proj_y=py, octave=octv, angle=ang, desc=pd, has_obs=(rng.random(n) < 0.85).astype(np.uint8))
    if ncam == 1:
        frame['u_right'] = np.where(rng.random(N) < 0.4, kps['x'] - rng.uniform(2, 40, N), -1).astype(np.float32)
        last['proj_ur'] = (px[:, 0] - rng.uniform(0, 45, n)).astype(np.float32)
    return frame, last


def _oracle_sbp_last(frame, last, th, fwd, bwd, check, max_dist=100, gate=None):
    L = O.lib()
    cam_off = frame['cam_off']; ncam = len(cam_off) - 1; N = int(cam_off[-1])
    kps = np.ascontiguousarray(frame['kps']); desc = np.ascontiguousarray(frame['desc'])
    grids = (C.c_void_p * 4)()
    for c in range(ncam):
        b = frame['bounds'][c]
        sub = np.ascontiguousarray(kps[cam_off[c]:cam_off[c + 1]])
        grids[c] = L.orc_grid_create(O.p(sub), sub.size, C.c_float(b[0]), C.c_float(b[1]), C.c_float(b[2]), C.c_float(b[3]))
    assigned = np.zeros(N, np.int32)
    opt = lambda a: None if a is None else O.p(a)
    n = L.orc_search_by_projection_last(ncam, O.p(cam_off), O.p(kps), O.p(desc), grids, O.p(frame['scale_factors']), O.p(frame['occupied']),
                                        opt(frame.get('u_right')), last['valid'].size, O.p(last['valid']), O.p(last['proj_x']),
                                        O.p(last['proj_y']), opt(last.get('proj_ur')), O.p(last['octave']), O.p(last['angle']),
                                        opt(last.get('has_obs')), O.p(last['desc']), C.c_float(th), (gate if gate is not None else 1 if fwd else 2 if bwd else 0), int(check), int(max_dist), O.p(assigned))
    for c in range(ncam):
        L.orc_grid_destroy(C.c_void_p(grids[c]))
    return n, assigned


@pytest.mark.parametrize('ncam,th,fwd,bwd,check', [(1, 15.0, False, False, True), (1, 7.0, True, False, True), (2, 15.0, False, True, True),
                                                   (2, 7.0, False, False, False), (4, 15.0, False, False, True), (4, 7.0, True, False, True)])
def test_search_by_projection_last_frame(ncam, th, fwd, bwd, check):
    """src/ORBmatcher.cc:1985-2413: claims in point order, empty camera-0 window skips the point, rotation-histogram filter."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, last = _sbp_last_case(ncam, seed=ncam * 5 + int(th))
    M = ob.ORBmatcher(0.9, checkOri=check)
    nm, assigned = M.SearchByProjectionLast(frame, last, th, fwd, bwd)
    rn, rassigned = _oracle_sbp_last(frame, last, th, fwd, bwd, check)
    assert rn == nm, (rn, nm)
    assert np.array_equal(assigned, rassigned), int((assigned != rassigned).sum())
    assert nm > 300
    cleared = M.SearchByProjectionLastCleared()
    if check:   # the filter did remove something: without it there are more matches
        n_all, a_all = _oracle_sbp_last(frame, last, th, fwd, bwd, False)
        assert n_all > rn
        # orbx_search_by_projection_last_cleared: exactly the keypoints whose claim the rotation filter cleared (NULL in the reference, :2404)
        assert cleared.sum() > 0 and (assigned[cleared == 1] == -1).all()
        assert np.array_equal(cleared == 1, (a_all >= 0) & (rassigned < 0))
    else:
        assert cleared.sum() == 0


@pytest.mark.parametrize('th', [1.0, 3.0])
def test_search_by_projection_rectified_stereo_gate(th):
    """SearchByProjection(Frame, MapPoints) on a stereo / RGB-D frame (F.Nleft == -1, mvuRight > 0): candidates whose right
    coordinate is further than the window radius from mTrackProjXR are skipped (src/ORBmatcher.cc:81-85)."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, points, _, nnratio = _sbp_case(1, seed=31 + int(th))
    rng = np.random.default_rng(5)
    N = len(frame['kps']); nmp = len(points['desc'])
    frame['u_right'] = np.where(rng.random(N) < 0.6, frame['kps']['x'] - rng.uniform(1, 30, N), -1).astype(np.float32)
    # half of the predicted right coordinates agree with the keypoint the point was planted on, half are off by up to 40 px
    points['proj_xr'] = (points['proj_x'][:, 0] - rng.uniform(0, 40, nmp)).astype(np.float32)
    M = ob.ORBmatcher(nnratio)
    nm, assigned = M.SearchByProjection(frame, points, th, False, 40.0)
    rn, rassigned = _oracle_sbp(frame, points, th, np.float32(nnratio), False, 40.0)
    assert rn == nm and np.array_equal(assigned, rassigned), (rn, nm, int((assigned != rassigned).sum()))
    # the gate did something: without it the result differs
    f2 = dict(frame); f2.pop('u_right'); p2 = dict(points); p2.pop('proj_xr')
    n0, a0 = M.SearchByProjection(f2, p2, th, False, 40.0)
    assert n0 > nm and not np.array_equal(a0, assigned)
    with pytest.raises(ob.OrbxError):                       # validation: predicted level out of range on an in-view entry
        bad = dict(points); bad['level'] = points['level'].copy(); bad['level'][np.nonzero(points['in_view'][:, 0])[0][0], 0] = 99
        M.SearchByProjection(frame, bad, th, False, 40.0)


@pytest.mark.parametrize('orb_dist,th', [(100, 10.0), (64, 15.0)])
def test_search_by_projection_relocalisation_overload(orb_dist, th):
    """SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:2415-2535) through the same entry point:
    camera 0 only, every claim visible, gate [l-1, l+1], ORBdist."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, last = _sbp_last_case(1, seed=21 + orb_dist)
    frame.pop('u_right'); last.pop('proj_ur'); last['has_obs'] = None
    M = ob.ORBmatcher(0.9)
    nm, assigned = M.SearchByProjectionLast(frame, last, th, max_dist=orb_dist)
    rn, rassigned = _oracle_sbp_last(frame, last, th, False, False, True, orb_dist)
    assert rn == nm and np.array_equal(assigned, rassigned)
    assert nm > 300
    # every claim is visible: no current keypoint receives two map points, none of the pre-occupied ones is touched
    assert (assigned[frame['occupied'] == 1] == -1).all()


@pytest.mark.parametrize('ratio_hamming,th', [(1.0, 6.0), (0.8, 10.0)])
def test_search_by_projection_keyframe_sim3_overloads(ratio_hamming, th):
    """SearchByProjection(KeyFrame, Sim3, vpPoints, vpMatched, th, ratioHamming, cameraID) (src/ORBmatcher.cc:668-893) through the
    same entry point: one camera block, gate [l-1, l], vpMatched as occupancy, TH_LOW * ratioHamming, no rotation filter."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, last = _sbp_last_case(1, seed=31 + int(th))
    frame.pop('u_right'); last.pop('proj_ur'); last['has_obs'] = None
    max_dist = int(np.floor(np.float32(50) * np.float32(ratio_hamming)))
    M = ob.ORBmatcher(0.9, checkOri=False)
    nm, assigned = M.SearchByProjectionLast(frame, last, th, max_dist=max_dist, level_gate=ob.GATE_BELOW)
    rn, rassigned = _oracle_sbp_last(frame, last, th, False, False, False, max_dist, gate=3)
    assert rn == nm and np.array_equal(assigned, rassigned)
    assert nm > 200
    k = np.nonzero(assigned >= 0)[0]
    d = frame['kps']['octave'][k] - last['octave'][assigned[k]]
    assert ((d == 0) | (d == -1)).all()


def _window_case(seed, nq=3000, stereo=True):
    rng = np.random.default_rng(seed)
    rig, res = _extract_frame(512, 512, 1500, [5100 + seed], (0, 511))
    kps = res[0][1].copy(); desc = res[0][2].copy()
    n = kps.size
    k = rng.integers(0, n, nq)
    qx = (kps['x'][k] + rng.uniform(-4, 4, nq)).astype(np.float32); qy = (kps['y'][k] + rng.uniform(-4, 4, nq)).astype(np.float32)
    lvl = np.clip(kps['octave'][k] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
    radius = (np.float32(3.0) * rig.scale[lvl]).astype(np.float32)
    qd = desc[k] ^ np.packbits(rng.random((nq, 256)) < 0.07, axis=1, bitorder='little')
    qd[::7] = synth_descriptors(len(qd[::7]), seed + 3)
    valid = (rng.random(nq) < 0.9).astype(np.uint8)
    skip = (rng.random(n) < 0.1).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, kps['x'] - rng.uniform(1, 30, n), -1).astype(np.float32) if stereo else None
    q_ur = (qx - rng.uniform(0, 32, nq)).astype(np.float32) if stereo else None
    inv_sigma2 = (1.0 / (rig.scale.astype(np.float32) ** 2)).astype(np.float32)
    return rig, kps, desc, (0.0, 512.0, 0.0, 512.0), qx, qy, radius, lvl, qd, valid, skip, u_right, q_ur, inv_sigma2


def _oracle_window(kps, desc, bounds, qx, qy, radius, mn, mx, qd, valid, skip, u_right, q_ur, inv_sigma2):
    L = O.lib()
    kps = np.ascontiguousarray(kps)
    g = L.orc_grid_create(O.p(kps), kps.size, C.c_float(bounds[0]), C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]))
    bi = np.zeros(qx.size, np.int32); bd = np.zeros(qx.size, np.int32)
    opt = lambda a: None if a is None else O.p(a)
    L.orc_window_best(C.c_void_p(g), O.p(kps), O.p(desc), opt(skip), opt(u_right), opt(inv_sigma2), qx.size, O.p(valid), O.p(qx), O.p(qy),
                      opt(q_ur), O.p(radius), O.p(mn), O.p(mx), O.p(qd), O.p(bi), O.p(bd))
    L.orc_grid_destroy(C.c_void_p(g))
    return bi, bd


@pytest.mark.parametrize('mode', ['fuse_stereo', 'fuse_mono', 'fuse_sim3'])
def test_window_best_fuse_loops(M, mode):
    """Search loops of both Fuse overloads (src/ORBmatcher.cc:1549-1601 with the reprojection gates, :1719-1738 without)."""
    rig, kps, desc, bounds, qx, qy, radius, lvl, qd, valid, skip, u_right, q_ur, inv_sigma2 = _window_case(7, stereo=mode == 'fuse_stereo')
    mn = (lvl - 1).astype(np.int32)
    kw = dict(valid=valid)
    if mode != 'fuse_sim3':
        kw.update(u_right=u_right, q_ur=q_ur, inv_level_sigma2=inv_sigma2)
    bi, bd = M.window_best(kps, desc, bounds, qx, qy, radius, mn, lvl, qd, **kw)
    ri, rd = _oracle_window(kps, desc, bounds, qx, qy, radius, mn, lvl, qd, valid, None, kw.get('u_right'), kw.get('q_ur'), kw.get('inv_level_sigma2'))
    assert np.array_equal(bi, ri) and np.array_equal(bd, rd)
    assert (bd <= 50).sum() > 1000 and (bi[valid == 0] == -1).all()
    if mode != 'fuse_sim3':   # the reprojection gate does reject candidates
        bi2, _ = M.window_best(kps, desc, bounds, qx, qy, radius, mn, lvl, qd, valid=valid)
        assert (bi2 != bi).any()


def test_window_best_skip_mask_and_empty(M):
    rig, kps, desc, bounds, qx, qy, radius, lvl, qd, valid, skip, *_ = _window_case(9, nq=500, stereo=False)
    mn = (lvl - 1).astype(np.int32)
    bi, bd = M.window_best(kps, desc, bounds, qx, qy, radius, mn, lvl, qd, valid=valid, skip=skip)
    ri, rd = _oracle_window(kps, desc, bounds, qx, qy, radius, mn, lvl, qd, valid, skip, None, None, None)
    assert np.array_equal(bi, ri) and np.array_equal(bd, rd)
    assert (skip[bi[bi >= 0]] == 0).all()
    bi, bd = M.window_best(kps[:0], desc[:0], bounds, qx, qy, radius, mn, lvl, qd)
    assert (bi == -1).all() and (bd == 256).all()
    bi, bd = M.window_best(kps, desc, bounds, qx[:0], qy[:0], radius[:0], mn[:0], lvl[:0], qd[:0])
    assert bi.size == 0


def test_search_by_sim3(M):
    """src/ORBmatcher.cc:1771-1983: two window searches + agreement."""
    rng = np.random.default_rng(12)
    rig, res = _extract_frame(512, 512, 1500, [5301, 5302], (0, 511))
    (_, k1, d1), (_, k2, d2) = [(r[0], r[1].copy(), r[2].copy()) for r in res]
    n1, n2 = k1.size, k2.size
    bounds = np.array([0, 512, 0, 512], np.float32)
    pair = rng.integers(0, n2, n1)                                    # map point of KF1 feature i1 is seen near KF2 feature pair[i1]

    def side(n_self, k_other, d_other, target):
        u = (k_other['x'][target] + rng.uniform(-2, 2, n_self)).astype(np.float32); v = (k_other['y'][target] + rng.uniform(-2, 2, n_self)).astype(np.float32)
        lvl = np.clip(k_other['octave'][target] + rng.integers(0, 2, n_self), 0, 7).astype(np.int32)
        desc = d_other[target] ^ np.packbits(rng.random((n_self, 256)) < 0.05, axis=1, bitorder='little')
        return dict(valid=(rng.random(n_self) < 0.85).astype(np.uint8), u=u, v=v, radius=(np.float32(7.5) * rig.scale[lvl]).astype(np.float32),
                    level=lvl, desc=desc)
    s1 = side(n1, k2, d2, pair)
    back = rng.integers(0, n1, n2)
    first = {}
    for i1 in range(n1):
        first.setdefault(int(pair[i1]), i1)
    for i2, i1 in first.items():
        if rng.random() < 0.8:
            back[i2] = i1                                             # mutual correspondences
    s2 = side(n2, k1, d1, back)
    nf, m12 = M.SearchBySim3(dict(kps=k1, desc=d1, bounds=bounds), dict(kps=k2, desc=d2, bounds=bounds), s1, s2)
    L = O.lib()
    g1 = L.orc_grid_create(O.p(k1), n1, C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
    g2 = L.orc_grid_create(O.p(k2), n2, C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
    ref = np.zeros(n1, np.int32)
    rn = L.orc_search_by_sim3(C.c_void_p(g1), O.p(k1), O.p(d1), n1, C.c_void_p(g2), O.p(k2), O.p(d2), n2, O.p(s1['valid']), O.p(s1['u']), O.p(s1['v']),
                              O.p(s1['radius']), O.p(s1['level']), O.p(s1['desc']), O.p(s2['valid']), O.p(s2['u']), O.p(s2['v']), O.p(s2['radius']),
                              O.p(s2['level']), O.p(s2['desc']), O.p(ref))
    L.orc_grid_destroy(C.c_void_p(g1)); L.orc_grid_destroy(C.c_void_p(g2))
    assert rn == nf and np.array_equal(ref, m12)
    assert nf > 100


@pytest.mark.parametrize('window,nnratio,check', [(10, 0.9, True), (25, 0.9, True), (40, 0.8, False)])
def test_search_for_initialization(window, nnratio, check):
    """src/ORBmatcher.cc:895-1004: level-0 windows, smallest-distance-so-far ownership with eviction, rotation filter."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    rng = np.random.default_rng(window)
    rig, res = _extract_frame(512, 512, 1500, [5400 + window], (0, 511))
    k2 = res[0][1].copy(); d2 = res[0][2].copy()
    n2 = k2.size
    n1 = 1400
    src = rng.integers(0, n2, n1)                                      # several F1 keypoints compete for the same F2 keypoint
    k1 = k2[src].copy()
    k1['x'] += rng.uniform(-6, 6, n1).astype(np.float32); k1['y'] += rng.uniform(-6, 6, n1).astype(np.float32)
    k1['angle'] = ((k2['angle'][src] + np.where(rng.random(n1) < 0.85, 15 + rng.uniform(-3, 3, n1), rng.uniform(0, 360, n1))) % 360).astype(np.float32)
    d1 = d2[src] ^ np.packbits(rng.random((n1, 256)) < rng.uniform(0.01, 0.08, (n1, 1)), axis=1, bitorder='little')
    d1[::9] = synth_descriptors(len(d1[::9]), 77)
    prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1).astype(np.float32))
    prev_ref = prev.copy()
    M = ob.ORBmatcher(nnratio, checkOri=check)
    nm, m12 = M.SearchForInitialization(k1, d1, k2, d2, (0, 512, 0, 512), prev, window)
    L = O.lib()
    g2 = L.orc_grid_create(O.p(k2), n2, C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
    ref = np.zeros(n1, np.int32)
    rn = L.orc_search_for_initialization(O.p(k1), O.p(d1), n1, C.c_void_p(g2), O.p(k2), O.p(d2), n2, O.p(prev_ref), window,
                                         C.c_float(np.float32(nnratio)), int(check), O.p(ref))
    L.orc_grid_destroy(C.c_void_p(g2))
    assert rn == nm, (rn, nm)
    assert np.array_equal(ref, m12), int((ref != m12).sum())
    assert np.array_equal(prev, prev_ref)
    lvl0 = int((k1['octave'] == 0).sum())
    assert nm > 0.25 * lvl0 and (k1['octave'][m12 >= 0] == 0).all()
    used = m12[m12 >= 0]
    assert np.unique(used).size == used.size                           # eviction keeps one owner per F2 keypoint


def test_search_by_projection_last_frame_degenerate():
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    frame, last = _sbp_last_case(2, seed=3, n=50)
    M = ob.ORBmatcher(0.9)
    last0 = {k: (v[:0] if isinstance(v, np.ndarray) else v) for k, v in last.items()}
    nm, assigned = M.SearchByProjectionLast(frame, last0, 7.0)
    assert nm == 0 and (assigned == -1).all()
    last['valid'][:] = 0
    nm, assigned = M.SearchByProjectionLast(frame, last, 7.0)
    assert nm == 0 and (assigned == -1).all()


def test_popc_peak_reports_a_rate():
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    assert ob.popc_peak() > 100.0


# MapPoint::ComputeDistinctiveDescriptors batched over map points
def test_distinctive_descriptors(M):
    rng = np.random.default_rng(12)
    sizes = np.concatenate([[0, 1, 2, 3, 33, 64, 130], rng.integers(1, 24, 400)]).astype(np.int32)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = synth_descriptors(len(sizes), 91)
    desc = np.repeat(base, sizes, axis=0)
    flips = np.packbits(rng.random((desc.shape[0], 256)) < 0.12, axis=1, bitorder='little')
    desc ^= flips
    desc[ptr[10]:ptr[11]] = desc[ptr[10]]                      # all-equal group: every median 0 -> first index
    best = M.ComputeDistinctiveDescriptors(desc, ptr)
    ref = np.zeros(len(sizes), np.int32)
    O.lib().orc_distinctive_descriptors(O.p(desc), O.p(ptr), len(sizes), O.p(ref))
    assert np.array_equal(best, ref)
    assert best[0] == -1 and best[1] == 0 and best[10] == 0


def test_knn2_parts_equals_single_call(M):
    """Train set given as parts (what the sharded matcher reads from peer GPUs) == one contiguous train set."""
    import torch
    t = synth_descriptors(5000, 21); q = synth_query_from_train(t, 22)[:1300]
    dq = torch.from_numpy(q).cuda(); parts = [torch.from_numpy(t[a:b].copy()).cuda() for a, b in ((0, 1), (1, 1700), (1700, 1700), (1700, 5000))]
    idx = torch.empty((1300, 2), dtype=torch.int32, device='cuda'); dd = torch.empty((1300, 2), dtype=torch.int32, device='cuda')
    torch.cuda.synchronize()
    M.knnMatch2_device_parts(dq.data_ptr(), 1300, [p.data_ptr() if p.numel() else 0 for p in parts], [p.shape[0] for p in parts],
                             idx.data_ptr(), dd.data_ptr())
    M.last_ms()                                  # synchronises on the matcher's stream
    ridx, rdist = O.knn2(q, t)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(dd.cpu().numpy(), rdist)


def test_knn2_parts_sync_waits_for_the_flags(M):
    """orbx_knn2_device_parts_sync on one GPU: parts and epoch flags in local memory; the flags are published from ANOTHER stream only
    after the call was queued, so the pull kernel really has to wait; result == one contiguous train set.  A flag that never arrives
    must end in a CUDA error (trap after the timeout), not in a hang — checked last, in a subprocess, because it kills the context."""
    import torch
    t = synth_descriptors(4000, 31); q = synth_query_from_train(t, 32)[:900]
    dq = torch.from_numpy(q).cuda()
    cuts = ((0, 1500), (1500, 1500), (1500, 2600), (2600, 4000))
    parts = [torch.from_numpy(t[a:b].copy()).cuda() for a, b in cuts]
    flags = torch.zeros((len(cuts), 64), dtype=torch.int32, device='cuda')     # one word per part, 256 bytes apart
    idx = torch.empty((900, 2), dtype=torch.int32, device='cuda'); dd = torch.empty((900, 2), dtype=torch.int32, device='cuda')
    torch.cuda.synchronize()
    L = ob.lib()
    L.orbx_shared_publish.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    side = torch.cuda.Stream()
    for epoch in (1, 2):
        M.knnMatch2_device_parts_sync(dq.data_ptr(), 900, [p.data_ptr() if p.numel() else 0 for p in parts], [p.shape[0] for p in parts], 2,
                                      [flags[i].data_ptr() for i in range(len(cuts))], epoch, idx.data_ptr(), dd.data_ptr(), timeout_s=30.0)
        time.sleep(0.05)                                                       # the pull kernel is spinning by now
        for i in (0, 3):
            assert L.orbx_shared_publish(0, C.c_void_p(flags[i].data_ptr()), epoch, C.c_void_p(side.cuda_stream)) == 0
        M.last_ms()
        ridx, rdist = O.knn2(q, t)
        assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(dd.cpu().numpy(), rdist)
    code = ("import numpy as np, torch, openmavis_b200 as ob\n"
            "m = ob.ORBmatcher(0.8)\n"
            "t = torch.zeros((64, 32), dtype=torch.uint8, device='cuda'); f = torch.zeros(64, dtype=torch.int32, device='cuda')\n"
            "o = torch.empty((64, 2), dtype=torch.int32, device='cuda'); d = torch.empty((64, 2), dtype=torch.int32, device='cuda')\n"
            "m.knnMatch2_device_parts_sync(t.data_ptr(), 64, [t.data_ptr(), t.data_ptr()], [64, 64], 0, [f.data_ptr(), f.data_ptr()], 1, o.data_ptr(), d.data_ptr(), timeout_s=0.5)\n"
            "try:\n    m.last_ms(); torch.cuda.synchronize(); print('NO ERROR')\nexcept Exception as e:\n    print('ERROR', type(e).__name__)\n")
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert 'ERROR' in r.stdout and 'NO ERROR' not in r.stdout, (r.stdout, r.stderr[-500:])


def _peer_worker(rank, world, port, n, out, sync=False):
    """One process per 'GPU' (all on device 0 when the box has a single GPU): publishes its train rows in a peer-mapped buffer,
    reads everybody's rows in place, answers its own query shard."""
    import os
    import torch
    import torch.distributed as dist
    from openmavis_b200 import shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        dev = rank % max(1, torch.cuda.device_count())
        torch.cuda.set_device(dev)
        train = synth_descriptors(n, 5); query = synth_query_from_train(train, 6)
        lo, cnt = shard.row_shard(n, rank, world)
        m = ob.ORBmatcher(0.8, device=dev)
        shared = ob.SharedRows(dist, dev, n)
        d_local = torch.from_numpy(train[lo:lo + cnt].copy()).cuda()
        d_q = torch.from_numpy(query[lo:lo + cnt].copy()).cuda()
        d_idx = torch.empty((cnt, 2), dtype=torch.int32, device='cuda'); d_dist = torch.empty((cnt, 2), dtype=torch.int32, device='cuda')
        for step in range(1, 5):                                 # exercises both buffers of the double buffer
            shared.fill(step & 1, d_local.data_ptr(), 0)
            if sync:                                             # device-side epoch flags: no barrier, no host synchronisation
                shared.publish(step, 0)
                m.knnMatch2_device_parts_sync(d_q.data_ptr(), cnt, shared.ptrs[step & 1], shared.rows, rank, shared.flags, step,
                                              d_idx.data_ptr(), d_dist.data_ptr(), timeout_s=60.0)
            else:
                torch.cuda.synchronize()
                dist.barrier()                                   # any barrier will do (gloo here)
                m.knnMatch2_device_parts(d_q.data_ptr(), cnt, shared.ptrs[step & 1], shared.rows, d_idx.data_ptr(), d_dist.data_ptr())
                m.last_ms()
        m.last_ms()
        full_idx, full_dist = shard.gather_results(d_idx.cpu().numpy(), d_dist.cpu().numpy(), n, dist)
        if rank == 0:
            np.savez(out, idx=full_idx, dist=full_dist)
        torch.cuda.synchronize()
        shared.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,sync', [(2, False), (3, False), (2, True), (3, True)])
def test_sharded_knn2_over_peer_mapped_rows(tmp_path, world, sync):
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    n, out = 3001, str(tmp_path / 'res.npz')
    mp.spawn(_peer_worker, args=(world, port, n, out, sync), nprocs=world, join=True)
    got = np.load(out)
    train = synth_descriptors(n, 5); query = synth_query_from_train(train, 6)
    ridx, rdist = O.knn2(query, train)
    assert np.array_equal(got['idx'], ridx) and np.array_equal(got['dist'], rdist)


def _synthetic_vocabulary(k, L, seed, irregular=False):
    """Random DBoW2-shaped tree: node 0 = root, every inner node up to level L-1 has k children (or 1..k and early leaves when
    irregular); leaves carry word ids in creation order and an idf-like weight (a few 0 = stopped words)."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]; children = [[]]; level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in frontier:
            nk = int(rng.integers(1, k + 1)) if irregular else k
            if irregular and lv > 1 and rng.random() < 0.15:
                continue                                           # early leaf
            for _ in range(nk):
                base = desc[p] if p else rng.integers(0, 256, 32, dtype=np.uint8)
                flips = np.packbits(rng.random(256) < (0.25 / lv), bitorder='little')
                desc.append(base ^ flips); children.append([]); level.append(lv)
                children[p].append(len(desc) - 1); nxt.append(len(desc) - 1)
        frontier = nxt
    n = len(desc)
    child_ptr = np.concatenate([[0], np.cumsum([len(c) for c in children])]).astype(np.int32)
    child = np.array([c for cs in children for c in cs], np.int32)
    word = np.full(n, -1, np.int32); weight = np.zeros(n, np.float64)
    leaves = [i for i in range(n) if not children[i]]
    word[leaves] = np.arange(len(leaves), dtype=np.int32)
    weight[leaves] = np.where(rng.random(len(leaves)) < 0.03, 0.0, rng.uniform(0.5, 9.0, len(leaves)))
    return child_ptr, child, np.stack(desc), word, weight, leaves


@pytest.mark.parametrize('k,L,irregular', [(10, 4, False), (10, 3, True), (40, 2, False), (3, 6, True)])
def test_vocabulary_transform(M, k, L, irregular):
    """TemplatedVocabulary::transform (tree descent on the device) vs the oracle, incl. BowVector / FeatureVector."""
    child_ptr, child, nd, word, weight, leaves = _synthetic_vocabulary(k, L, 100 + k + L, irregular)
    rng = np.random.default_rng(5)
    n = 1500
    feat = nd[rng.choice(leaves, n)] ^ np.packbits(rng.random((n, 256)) < 0.08, axis=1, bitorder='little')
    feat[::7] = synth_descriptors((n + 6) // 7, 17)                # unrelated features
    feat[1] = feat[0]                                              # duplicates accumulate in the same word
    voc = ob.ORBVocabulary(child_ptr, child, nd, word, weight, L)
    levelsup = 4 if L > 4 else 1
    (ids, vals), (nodes, ptr, idx), gw, gn = voc.transform(M, feat, levelsup)
    rw = np.zeros(n, np.int32); rn = np.zeros(n, np.int32)
    Lb = O.lib()
    Lb.orc_vocab_transform(O.p(child_ptr), O.p(child), O.p(nd), O.p(word), L, O.p(feat), n, levelsup, O.p(rw), O.p(rn))
    assert np.array_equal(gw, rw) and np.array_equal(gn, rn)
    leaf_of_word = np.zeros(len(leaves), np.int64); leaf_of_word[word[leaves]] = leaves
    w = weight[leaf_of_word[rw]]
    bi = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fp = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
    nfn = C.c_int(0)
    nb = Lb.orc_bow_from_words(O.p(rw), O.p(rn), O.p(w), n, O.p(bi), O.p(bv), O.p(fn), O.p(fp), O.p(fi), C.byref(nfn))
    assert np.array_equal(ids, bi[:nb]) and np.array_equal(vals.view(np.uint64), bv[:nb].view(np.uint64))
    assert np.array_equal(nodes, fn[:nfn.value]) and np.array_equal(ptr, fp[:nfn.value + 1]) and np.array_equal(idx, fi[:fp[nfn.value]])
    assert nb >= 3


def test_vocabulary_objects_created_one_after_another(M):
    """Vocabularies created and destroyed in sequence (the allocator hands out the same address again): the word -> weight map
    belongs to each vocabulary, the weights of the second one must not come from the first."""
    rng = np.random.default_rng(9)
    outs = []
    for round_ in range(3):
        child_ptr, child, nd, word, weight, leaves = _synthetic_vocabulary(6, 3, 300 + round_)
        weight = weight * (1.0 + round_)                               # same tree shape, different weights / word numbering
        word = word.copy(); word[leaves] = rng.permutation(len(leaves)).astype(word.dtype)
        feat = nd[rng.choice(leaves, 400)]
        voc = ob.ORBVocabulary(child_ptr, child, nd, word, weight, 3)
        (ids, vals), _, gw, gn = voc.transform(M, feat, 1)
        leaf_of_word = np.zeros(len(leaves), np.int64); leaf_of_word[word[leaves]] = leaves
        exp = {}
        for wd in gw:
            if weight[leaf_of_word[wd]] > 0:                          # stopped words carry no weight
                exp[int(wd)] = exp.get(int(wd), 0.0) + weight[leaf_of_word[wd]]
        norm = sum(exp[k_] for k_ in sorted(exp))                     # BowVector::normalize(L1)
        got = dict(zip(ids.tolist(), vals.tolist()))
        assert set(got) == set(exp) and len(exp) > 20
        for k_ in exp:
            assert abs(got[k_] - exp[k_] / norm) <= 1e-12, (round_, k_)
        del voc


def _random_frame(rng, ncam, n_per_cam, size=200.0, pool=6):
    """Keypoints without an extraction: random positions / octaves, descriptors from a tiny pool with a few flipped bits (many
    equal distances, dense windows, long claim chains)."""
    ns = [int(rng.integers(max(1, n_per_cam // 2), n_per_cam + 1)) for _ in range(ncam)]
    cam_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    N = int(cam_off[-1])
    kps = np.zeros(N, ob.KP_DTYPE)
    kps['x'] = rng.uniform(0, size, N); kps['y'] = rng.uniform(0, size, N)
    kps['octave'] = rng.integers(0, 8, N); kps['angle'] = rng.uniform(0, 360, N); kps['size'] = 31; kps['class_id'] = -1
    base = synth_descriptors(pool, int(rng.integers(1 << 30)))
    desc = base[rng.integers(0, pool, N)] ^ np.packbits(rng.random((N, 256)) < 0.03, axis=1, bitorder='little')
    sf = np.array([np.float32(1.2) ** l for l in range(8)], np.float32)
    frame = dict(cam_off=cam_off, kps=kps, desc=desc, bounds=np.tile(np.array([0, size, 0, size], np.float32), (ncam, 1)), scale_factors=sf,
                 occupied=(rng.random(N) < 0.1).astype(np.uint8))
    return frame, base


@pytest.mark.parametrize('seed', range(int(os.environ.get('ORBX_STRESS_N', '12'))))
def test_ordered_resolver_stress(seed):
    """Many small, dense, adversarial cases of the three ordered matchers (claims in point order, eviction, mirrors): every one must
    equal the sequential oracle.  Sizes straddle the resolver's 1024-point window and its 12-entry touch cache."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    rng = np.random.default_rng(1000 + seed)
    L = O.lib()
    for rep in range(4):
        ncam = int(rng.choice([1, 2, 4]))
        frame, base = _random_frame(rng, ncam, int(rng.choice([8, 60, 400])))
        cam_off, kps, N = frame['cam_off'], frame['kps'], int(frame['cam_off'][-1])
        n = int(rng.choice([3, 200, 1100, 2500]))
        tgt = rng.integers(0, N, n)
        spread = float(rng.choice([0.5, 4.0, 15.0]))
        px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
        for c in range(ncam):
            kk = np.where((tgt >= cam_off[c]) & (tgt < cam_off[c + 1]), tgt, rng.integers(cam_off[c], cam_off[c + 1], n))
            px[:, c] = kps['x'][kk] + rng.uniform(-spread, spread, n); py[:, c] = kps['y'][kk] + rng.uniform(-spread, spread, n)
        pdesc = base[rng.integers(0, base.shape[0], n)] ^ np.packbits(rng.random((n, 256)) < 0.03, axis=1, bitorder='little')
        octv = np.clip(kps['octave'][tgt] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
        has_obs = (rng.random(n) < float(rng.choice([0.5, 0.9, 1.0]))).astype(np.uint8)
        # ---- mode 1: SearchByProjection(CurrentFrame, LastFrame) and its overloads
        last = dict(valid=(rng.random(n) < 0.9).astype(np.uint8), proj_x=px, proj_y=py, octave=octv,
                    angle=rng.uniform(0, 360, n).astype(np.float32), desc=pdesc, has_obs=has_obs)
        th = float(rng.choice([2.0, 7.0, 15.0])); gate = int(rng.integers(0, 4)) if ncam == 1 else int(rng.integers(0, 3))
        check = bool(rng.integers(0, 2)); md = int(rng.choice([100, 50, 20]))
        M1 = ob.ORBmatcher(0.9, checkOri=check)
        nm, assigned = M1.SearchByProjectionLast(frame, last, th, max_dist=md, level_gate=gate)
        rn, rassigned = _oracle_sbp_last(frame, last, th, False, False, check, md, gate=gate)
        assert rn == nm and np.array_equal(assigned, rassigned), ('mode 1', seed, rep, ncam, n, N, th, gate, check, md)
    # ---- mode 0 (SearchByProjection(Frame, MapPoints) with stereo mirror tables) and mode 2 (SearchForInitialization)
    rng = np.random.default_rng(2000 + seed)
    for rep in range(3):
        ncam = int(rng.choice([1, 2, 4]))
        frame, base = _random_frame(rng, ncam, int(rng.choice([8, 60, 400])))
        cam_off, kps, N = frame['cam_off'], frame['kps'], int(frame['cam_off'][-1])
        n = int(rng.choice([3, 200, 1100, 2500]))
        tgt = rng.integers(0, N, n)
        spread = float(rng.choice([0.5, 4.0, 15.0]))
        px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
        for c in range(ncam):
            kk = np.where((tgt >= cam_off[c]) & (tgt < cam_off[c + 1]), tgt, rng.integers(cam_off[c], cam_off[c + 1], n))
            px[:, c] = kps['x'][kk] + rng.uniform(-spread, spread, n); py[:, c] = kps['y'][kk] + rng.uniform(-spread, spread, n)
        pdesc = base[rng.integers(0, base.shape[0], n)] ^ np.packbits(rng.random((n, 256)) < 0.03, axis=1, bitorder='little')
        octv = np.clip(kps['octave'][tgt] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
        l2r = r2l = None
        if ncam >= 2:
            nl, nr = int(cam_off[1]), int(cam_off[2] - cam_off[1])
            l2r = np.full(nl, -1, np.int32); r2l = np.full(nr, -1, np.int32)
            k = min(nl, nr) // 2
            for a, b in zip(rng.permutation(nl)[:k], rng.permutation(nr)[:k]):
                l2r[a] = b; r2l[b] = a
        f0 = dict(frame, l2r=l2r, r2l=r2l)
        in_view = (rng.random((n, 4)) < 0.8).astype(np.uint8); in_view[:, ncam:] = 0
        points = dict(in_view=in_view, proj_x=px, proj_y=py, level=np.repeat(octv[:, None], 4, 1).astype(np.int32),
                      view_cos=rng.choice([0.9990, 0.99], (n, 4)).astype(np.float32), desc=pdesc, depth=rng.uniform(1, 80, n).astype(np.float32),
                      bad=(rng.random(n) < 0.05).astype(np.uint8), has_obs=(rng.random(n) < float(rng.choice([0.5, 0.9, 1.0]))).astype(np.uint8))
        nnr = float(rng.choice([0.6, 0.8, 1.0])); th = float(rng.choice([1.0, 3.0, 6.0])); far = bool(rng.integers(0, 2))
        M0 = ob.ORBmatcher(nnr)
        nm, assigned = M0.SearchByProjection(f0, points, th, far, 40.0)
        rn, rassigned = _oracle_sbp(f0, points, th, np.float32(nnr), far, 40.0)
        assert rn == nm and np.array_equal(assigned, rassigned), ('mode 0', seed, rep, ncam, n, N, th, nnr, far)
        # ---- mode 2: SearchForInitialization on camera 0's keypoints
        k2 = np.ascontiguousarray(kps[:cam_off[1]]); d2 = np.ascontiguousarray(frame['desc'][:cam_off[1]])
        k2['octave'] = np.where(rng.random(k2.size) < 0.7, 0, k2['octave'])
        n1 = min(n, 1500)
        src = rng.integers(0, k2.size, n1)
        k1 = k2[src].copy()
        k1['x'] += rng.uniform(-spread, spread, n1).astype(np.float32); k1['y'] += rng.uniform(-spread, spread, n1).astype(np.float32)
        k1['angle'] = rng.uniform(0, 360, n1).astype(np.float32)
        d1 = np.ascontiguousarray(pdesc[:n1])
        prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1).astype(np.float32)); prev_ref = prev.copy()
        win = int(rng.choice([3, 10, 30])); check = bool(rng.integers(0, 2))
        M2 = ob.ORBmatcher(float(rng.choice([0.7, 0.9, 1.0])), checkOri=check)
        nm, m12 = M2.SearchForInitialization(k1, d1, k2, d2, (0, 200, 0, 200), prev, win)
        g2 = L.orc_grid_create(O.p(k2), k2.size, C.c_float(0), C.c_float(200), C.c_float(0), C.c_float(200))
        ref = np.zeros(n1, np.int32)
        rn = L.orc_search_for_initialization(O.p(k1), O.p(d1), n1, C.c_void_p(g2), O.p(k2), O.p(d2), k2.size, O.p(prev_ref), win,
                                             C.c_float(np.float32(M2.mfNNratio)), int(check), O.p(ref))
        L.orc_grid_destroy(C.c_void_p(g2))
        assert rn == nm and np.array_equal(ref, m12) and np.array_equal(prev, prev_ref), ('mode 2', seed, rep, n1, k2.size, win, check)


# ---- SearchForTriangulation, the whole function incl. bCoarse = false (src/ORBmatcher.cc:1131-1456) ----------------------------------
def _sft_scenario(rng, ncam, trial):
    per1 = [int(rng.integers(150, 500)) for _ in range(ncam)]; per2 = [int(rng.integers(150, 500)) for _ in range(ncam)]
    off1 = np.concatenate([[0], np.cumsum(per1)]).astype(np.int32); off2 = np.concatenate([[0], np.cumsum(per2)]).astype(np.int32)
    n1, n2 = int(off1[-1]), int(off2[-1])
    k2 = np.zeros(n2, ob.KP_DTYPE)
    k2['x'] = rng.uniform(20, 620, n2); k2['y'] = rng.uniform(20, 460, n2); k2['octave'] = rng.integers(0, 8, n2); k2['angle'] = rng.uniform(0, 360, n2)
    d2 = synth_descriptors(n2, 4000 + trial)
    d2[10:60] = d2[10]                                       # exact duplicates -> last-minimum tie rule
    src = rng.integers(0, n2, n1)
    flips = np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
    d1 = d2[src] ^ flips
    d1[::6] = synth_descriptors(len(d1[::6]), 43 + trial)
    k1 = k2[src].copy()
    k1['x'] += rng.uniform(-30, 30, n1).astype(np.float32); k1['y'] += rng.uniform(-2.5, 2.5, n1).astype(np.float32)   # near the epipolar line y2 = y1
    k1['angle'] = (k1['angle'] + rng.choice([0, 0, 0, 45, 200], n1)) % 360

    def fvec(desc, nbits=5):
        node = desc[:, 0].astype(np.int32) & ((1 << nbits) - 1)
        order = np.argsort(node, kind='stable')
        nodes, counts = np.unique(node, return_counts=True)
        return nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)
    fv1, fv2 = fvec(d1), fvec(d2)
    fv1 = (fv1[0][::2].copy(), np.concatenate([[0], np.cumsum(np.diff(fv1[1])[::2])]).astype(np.int32),
           np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in range(0, len(fv1[0]), 2)]).astype(np.int32))   # some nodes only in KF2
    kf1 = dict(cam_off=off1, kps=k1, desc=d1, has_mp=(rng.random(n1) < 0.3).astype(np.uint8), u_right=None, fv=fv1)
    kf2 = dict(cam_off=off2, kps=k2, desc=d2, has_mp=(rng.random(n2) < 0.3).astype(np.uint8), u_right=None, fv=fv2)
    only_stereo = False
    if ncam == 1 and trial % 2 == 1:
        kf1['u_right'] = np.where(rng.random(n1) < 0.5, 10.0, -1.0).astype(np.float32)
        kf2['u_right'] = np.where(rng.random(n2) < 0.5, 10.0, -1.0).astype(np.float32)
        only_stereo = trial % 4 == 3
    e = int(rng.integers(0, n2))
    ep = np.array([k2['x'][e] + 2.0, k2['y'][e] - 1.0], np.float32)
    return kf1, kf2, ep, only_stereo


def _oracle_sft(kf1, kf2, ep, sf, s1, s2, pairs, only_stereo, coarse, chk, want_margin=False):
    L = O.lib()
    n1 = len(kf1['kps'])
    ref = np.zeros(n1, np.int32)
    mm = C.c_float(0)
    fv1, fv2 = kf1['fv'], kf2['fv']
    rn = L.orc_search_for_triangulation_cams(len(kf1['cam_off']) - 1, O.p(kf1['cam_off']), O.p(kf1['kps']), O.p(kf1['desc']), O.p(kf1['has_mp']),
                                             None if kf1['u_right'] is None else O.p(kf1['u_right']), len(fv1[0]), O.p(fv1[0]), O.p(fv1[1]),
                                             O.p(fv1[2]), O.p(kf2['cam_off']), O.p(kf2['kps']), O.p(kf2['desc']), O.p(kf2['has_mp']),
                                             None if kf2['u_right'] is None else O.p(kf2['u_right']), len(fv2[0]), O.p(fv2[0]), O.p(fv2[1]),
                                             O.p(fv2[2]), O.p(ep), O.p(sf), O.p(s1), O.p(s2), int(only_stereo), int(coarse), int(chk),
                                             None if pairs is None else O.p(pairs), O.p(ref), C.byref(mm) if want_margin else None)
    return rn, ref, mm.value


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_for_triangulation_full_pinhole(M, ncam):
    """bCoarse = false with Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:103-132): every camera pair gets its own F12 (epipolar
    line y2 = y1 + offset, a different offset per pair), so a candidate whose cameras have no branch in :1301-1378 passes or fails
    depending on the carried pair.  Bit-exact match lists; also bCoarse = true and the rotation filter through the same entry point."""
    rng = np.random.default_rng(70 + ncam)
    sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
    s1 = (sf * sf).astype(np.float32); s2 = (sf * sf * np.float32(1.0)).astype(np.float32)
    pairs = ob.epipolar_pairs()
    for s in range(11):
        pairs['model'][s] = ob.CAM_PINHOLE
        pairs['F12'][s] = np.array([0, 0, 0, 0, 0, -1, 0, 1, -(s % 4 - 1.5) * 1.25], np.float32)
    total = 0
    for trial in range(8):
        kf1, kf2, ep, only_stereo = _sft_scenario(rng, ncam, trial)
        for coarse in (False, True):
            chk = (trial + coarse) % 2 == 1
            nm, prs, m12 = M.SearchForTriangulationFull(kf1, kf2, ep, sf, s1, s2, pairs, only_stereo, coarse, chk)
            rn, ref, _ = _oracle_sft(kf1, kf2, ep, sf, s1, s2, pairs, only_stereo, coarse, chk)
            assert nm == rn, (trial, coarse, nm, rn)
            assert np.array_equal(m12, ref), (trial, coarse, int((m12 != ref).sum()))
            assert prs == [(int(i), int(ref[i])) for i in np.nonzero(ref >= 0)[0]]
            total += nm
        # the coarse result through the array-level entry point of round 1 must agree as well (no rotation filter there)
        if not only_stereo and kf1['u_right'] is None:
            nm0, _, m0 = M.SearchForTriangulation(kf1['fv'], kf2['fv'], kf1['desc'], kf2['desc'], kf1['has_mp'], kf2['has_mp'])
            _, _, mc = M.SearchForTriangulationFull(kf1, kf2, ep, sf, s1, s2, None, False, True, False)
            if ncam > 1:                                        # with ncam == 1 the epipole gate also applies to bCoarse
                assert np.array_equal(m0, mc)
    assert total > 200


def test_search_for_triangulation_full_kb8(M):
    """bCoarse = false on a 4-camera Kannala-Brandt rig: KannalaBrandt8::epipolarConstrain = TriangulateMatches(...) > 0.0001
    (src/CameraModels/KannalaBrandt8.cpp:219-230).  Float geometry (atan2f / tanf / 4x4 null vector): tolerance-based as DESIGN.md §2
    states — the match lists must be identical whenever no tested quantity of the oracle run lies within 1e-3 (relative) of its gate."""
    from openmavis_b200.synth import synth_fisheye_pairs
    rng = np.random.default_rng(91)
    sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
    s1 = (sf * sf).astype(np.float32); s2 = s1.copy()
    checked = 0
    for trial in range(12):
        fs = synth_fisheye_pairs(1200, 500 + trial, outlier_frac=0.3)
        pairs = ob.epipolar_pairs()
        for s in range(11):
            pairs['model'][s] = ob.CAM_KB8
            t12 = fs['t12'] * np.float32(1.0 + 0.02 * (s % 3))          # the pairs differ a little: the carried state changes outcomes
            pairs['kb8'][s] = ob.kb8_stereo(fs['cam1'], fs['cam2'], fs['R12'], t12)[0]
        kf1, kf2, ep, _ = _sft_scenario(rng, 4, trial)
        n1, n2 = len(kf1['kps']), len(kf2['kps'])
        # keypoints: observations of the same 3-D points in the two cameras (planted partner of idx1 is found through the descriptors)
        pick2 = rng.integers(0, 1200, n2)
        kf2['kps']['x'], kf2['kps']['y'] = fs['xy2'][pick2, 0], fs['xy2'][pick2, 1]
        d2 = kf2['desc']
        src = rng.integers(0, n2, n1)
        flips = np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        d1 = d2[src] ^ flips
        kf1['desc'] = d1
        kf1['kps']['x'], kf1['kps']['y'] = fs['xy1'][pick2[src], 0], fs['xy1'][pick2[src], 1]
        node = d1[:, 0].astype(np.int32) & 31
        order = np.argsort(node, kind='stable'); nodes, counts = np.unique(node, return_counts=True)
        kf1['fv'] = (nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32))
        rn, ref, margin = _oracle_sft(kf1, kf2, ep, sf, s1, s2, pairs, False, False, False, want_margin=True)
        if margin < 1e-3:
            continue                                                    # a gate is too close to call in float: not a parity case
        nm, prs, m12 = M.SearchForTriangulationFull(kf1, kf2, ep, sf, s1, s2, pairs, False, False, False)
        assert nm == rn and np.array_equal(m12, ref), (trial, nm, rn, int((m12 != ref).sum()))
        assert nm > 50
        checked += 1
    assert checked >= 4


# ---- guided matchers batched across frames (orbx_matcher_pool) --------------------------------------------------------------------
def _random_points_case(rng, ncam):
    frame, base = _random_frame(rng, ncam, int(rng.choice([60, 400])))
    cam_off, kps, N = frame['cam_off'], frame['kps'], int(frame['cam_off'][-1])
    n = int(rng.choice([200, 1100, 2500]))
    tgt = rng.integers(0, N, n)
    px = np.zeros((n, 4), np.float32); py = np.zeros((n, 4), np.float32)
    for c in range(ncam):
        kk = np.where((tgt >= cam_off[c]) & (tgt < cam_off[c + 1]), tgt, rng.integers(cam_off[c], cam_off[c + 1], n))
        px[:, c] = kps['x'][kk] + rng.uniform(-4, 4, n); py[:, c] = kps['y'][kk] + rng.uniform(-4, 4, n)
    pdesc = base[rng.integers(0, base.shape[0], n)] ^ np.packbits(rng.random((n, 256)) < 0.03, axis=1, bitorder='little')
    octv = np.clip(kps['octave'][tgt] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
    in_view = (rng.random((n, 4)) < 0.8).astype(np.uint8); in_view[:, ncam:] = 0
    points = dict(in_view=in_view, proj_x=px, proj_y=py, level=np.repeat(octv[:, None], 4, 1).astype(np.int32),
                  view_cos=rng.choice([0.9990, 0.99], (n, 4)).astype(np.float32), desc=pdesc, depth=rng.uniform(1, 80, n).astype(np.float32),
                  bad=(rng.random(n) < 0.05).astype(np.uint8), has_obs=(rng.random(n) < 0.9).astype(np.uint8))
    return dict(frame, l2r=None, r2l=None), points


def test_matcher_pool_batches_equal_single_calls():
    """Each call of a batch must give exactly the result of the single-frame entry point (and therefore of the oracle): 24 different
    SearchByProjection(Frame, MapPoints) frames, 12 SearchByProjection(CurrentFrame, LastFrame) frames and 12 SearchByBoW(KF, Frame)
    pairs on a pool of 8 lanes."""
    if ob.device_count() == 0:
        pytest.fail('no CUDA device')
    pool = ob.MatcherPool(lanes=8)
    M = ob.ORBmatcher(0.8)
    rng = np.random.default_rng(123)
    frames, points = [], []
    for i in range(24):
        f, p_ = _random_points_case(rng, ncam=[1, 2, 4][i % 3])
        frames.append(f); points.append(p_)
    got = pool.SearchByProjectionBatch(frames, points, 3.0, False, 50.0, 0.8)
    for i in range(24):
        nm, a = M.SearchByProjection(frames[i], points[i], 3.0, False, 50.0)
        assert got[i][0] == nm and np.array_equal(got[i][1], a), i
        rn, ra = _oracle_sbp(frames[i], points[i], 3.0, np.float32(0.8), False, 50.0)
        assert got[i][0] == rn and np.array_equal(got[i][1], ra), i
    M9 = ob.ORBmatcher(0.9, checkOri=True)
    fl = [_sbp_last_case([1, 2, 4][i % 3], seed=900 + i, n=1500) for i in range(12)]
    got = pool.SearchByProjectionLastBatch([f for f, _ in fl], [la for _, la in fl], 7.0)
    for i, (f, la) in enumerate(fl):
        nm, a = M9.SearchByProjectionLast(f, la, 7.0)
        assert got[i][0] == nm and np.array_equal(got[i][1], a), i
    cases = [_bow_case(700 + i, 900 + 50 * i, 1000, 6) for i in range(12)]      # (d1, d2, a1, a2, fv1, fv2, skip1, skip2)
    M7 = ob.ORBmatcher(0.7, checkOri=True)
    got = pool.SearchByBoWBatch([c[4] for c in cases], [c[5] for c in cases], [c[0] for c in cases], [c[1] for c in cases],
                                [c[6] for c in cases], [c[2] for c in cases], [c[3] for c in cases],
                                [np.array([0, len(c[1])], np.int32) for c in cases], 0.7, True)
    for i, c in enumerate(cases):
        nm, mf = M7.SearchByBoW(c[4], c[5], c[0], c[1], c[6], c[2], c[3], None)
        assert got[i][0] == nm and np.array_equal(got[i][1], mf), i
    pool.close()
