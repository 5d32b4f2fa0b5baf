"""oracle/_ref/libref_matcher.so — the reference's OWN src/ORBmatcher.cc (unmodified) plus the grid functions of src/Frame.cc, compiled
against oracle/refstub (oracle/Makefile) — against the restated oracle (oracle/orb_oracle.cpp) on the same flat arrays.  This pins the
matcher restatements (selection order, gates, best / second bookkeeping, claims, rotation histogram, carried camera-pair state) to the
reference's object code.  CPU only; tests/test_gpu_match.py compares the CUDA path with the same oracle functions."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_descriptors

TRIALS_X = int(os.environ.get('ORBX_GLUE_TRIALS_X', '1'))   # multiplies the random scenarios per test (a wide sweep before a release)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KP = O.KP_DTYPE
_lib = None


def RM():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'oracle', '_ref', 'libref_matcher.so')
        if not os.path.exists(so):
            O.ref_lib()                                      # runs make when /root/reference is mounted
        if not os.path.exists(so):
            raise RuntimeError('%s is missing and /root/reference is not mounted to build it' % so)
        O.lib()
        _lib = C.CDLL(so)
    return _lib


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_kps(n, rng, w=640, h=480, levels=8, lattice=None):
    k = np.zeros(n, KP)
    if lattice:
        k['x'] = rng.integers(0, w // lattice, n) * lattice; k['y'] = rng.integers(0, h // lattice, n) * lattice
    else:
        k['x'] = rng.uniform(0, w, n).astype(np.float32); k['y'] = rng.uniform(0, h, n).astype(np.float32)
    k['octave'] = rng.integers(0, levels, n); k['angle'] = rng.uniform(0, 360, n).astype(np.float32)
    k['size'] = 31; k['class_id'] = -1
    return k


def test_descriptor_distance():
    a = synth_descriptors(2000, 1); b = synth_descriptors(2000, 2)
    b[:10] = a[:10]; b[10:20] = ~a[10:20]
    L = RM()
    for i in range(2000):
        assert L.refm_descriptor_distance(p(a[i]), p(b[i])) == O.hamming(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


def test_three_maxima():
    rng = np.random.default_rng(5)
    L = RM()
    O.lib().orc_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    for trial in range(400 * TRIALS_X):
        sizes = rng.integers(0, [3, 10, 60][trial % 3], 30).astype(np.int32)
        if trial % 7 == 0:
            sizes[rng.integers(0, 30, 3)] = sizes.max()       # ties between the maxima
        a = np.zeros(3, np.int32); b = np.zeros(3, np.int32)
        L.refm_three_maxima(p(sizes), 30, p(a))
        O.lib().orc_three_maxima(p(sizes), 30, p(b))
        assert np.array_equal(a, b), (trial, sizes, a, b)


def test_frame_grid_get_features_in_area():
    rng = np.random.default_rng(6)
    L = RM()
    for trial in range(60 * TRIALS_X):
        n = int(rng.integers(1, 1500))
        k = make_kps(n, rng, lattice=4 if trial % 3 == 0 else None)
        bounds = np.array([0, 640, 0, 480], np.float32) if trial % 2 else np.array([-13.5, 655.25, -9.0, 489.75], np.float32)
        g = O.lib().orc_grid_create(p(k), n, C.c_float(bounds[0]), C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]))
        for qi in range(40):
            x, y = np.float32(rng.uniform(-30, 670)), np.float32(rng.uniform(-30, 510))
            r = np.float32(rng.choice([2.5, 4.0, 7.5, 15.0, 42.0, 200.0]))
            lo, hi = [(-1, -1), (0, 0), (2, 3), (0, 7), (3, -1)][qi % 5]
            a = np.zeros(n, np.int32); b = np.zeros(n, np.int32)
            na = L.refm_grid_query(p(k), n, p(bounds), C.c_float(x), C.c_float(y), C.c_float(r), lo, hi, p(a), n)
            nb = O.lib().orc_grid_query(C.c_void_p(g), p(k), C.c_float(x), C.c_float(y), C.c_float(r), lo, hi, p(b), n)
            assert na == nb and np.array_equal(a[:na], b[:nb]), (trial, qi)
        O.lib().orc_grid_destroy(C.c_void_p(g))


def sbp_scenario(rng, ncam, nmp, crowded):
    w, h = 512, 512
    per = [int(rng.integers(200, 700)) for _ in range(ncam)]
    cam_off = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    N = int(cam_off[-1])
    k = make_kps(N, rng, w, h, lattice=8 if crowded else None)
    desc = synth_descriptors(N, int(rng.integers(1 << 30)))
    sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
    in_view = (rng.random((nmp, 4)) < np.array([0.9, 0.6, 0.5, 0.5])).astype(np.uint8)
    src = np.stack([rng.integers(cam_off[min(c, ncam - 1)], cam_off[min(c, ncam - 1) + 1], nmp) for c in range(4)], 1)
    px = (k['x'][src] + rng.uniform(-3, 3, (nmp, 4))).astype(np.float32); py = (k['y'][src] + rng.uniform(-3, 3, (nmp, 4))).astype(np.float32)
    lvl = np.clip(k['octave'][src] + rng.integers(0, 2, (nmp, 4)), 0, 7).astype(np.int32)
    lvl[rng.random((nmp, 4)) < 0.05] = 0
    vc = rng.choice(np.array([0.9990, 0.99], np.float32), (nmp, 4)).astype(np.float32)
    md = synth_descriptors(nmp, int(rng.integers(1 << 30)))
    own = rng.random(nmp) < 0.7
    md[own] = desc[src[own, 0]] ^ np.packbits(rng.random((int(own.sum()), 256)) < 0.08, axis=1, bitorder='little')
    if crowded:
        md[: nmp // 4] = md[0]                               # equal descriptors: ties and claims on the same keypoints
    occupied = (rng.random(N) < 0.1).astype(np.uint8)
    l2r = r2l = None
    if ncam >= 2:
        nl, nr = per[0], per[1]
        l2r = np.full(nl, -1, np.int32); r2l = np.full(nr, -1, np.int32)
        m = min(nl, nr) // 3
        a = rng.permutation(nl)[:m]; b = rng.permutation(nr)[:m]
        l2r[a] = b; r2l[b] = a
    u_right = None; pxr = np.zeros(nmp, np.float32)
    if ncam == 1:
        u_right = np.where(rng.random(N) < 0.5, k['x'] - rng.uniform(1, 30, N), -1).astype(np.float32)
        pxr = (px[:, 0] - rng.uniform(0, 40, nmp)).astype(np.float32)
    depth = rng.uniform(1, 80, nmp).astype(np.float32)
    bad = (rng.random(nmp) < 0.05).astype(np.uint8); has_obs = (rng.random(nmp) < 0.9).astype(np.uint8)
    return dict(ncam=ncam, cam_off=cam_off, k=k, desc=desc, sf=sf, occupied=occupied, l2r=l2r, r2l=r2l, u_right=u_right, nmp=nmp, in_view=in_view,
                px=px, py=py, lvl=lvl, vc=vc, pxr=pxr, depth=depth, bad=bad, has_obs=has_obs, md=md, bounds=np.array([0, w, 0, h], np.float32))


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_projection_frame_mappoints(ncam):
    """SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints), src/ORBmatcher.cc:23-340, incl. the sequential
    already-claimed rule, the ratio-test `continue` of the MapPoint loop, the camera-0-only th, the rectified-stereo gate (:81-85)
    and the left/right mirror writes."""
    rng = np.random.default_rng(100 + ncam)
    L = RM(); Lo = O.lib()
    for trial in range(12 * TRIALS_X):
        S = sbp_scenario(rng, ncam, int(rng.integers(200, 1500)), crowded=trial % 2 == 1)
        th = float([1.0, 3.0, 6.0][trial % 3]); far = trial % 4 == 3; nnr = 0.8
        N = int(S['cam_off'][-1])
        a = np.zeros(N, np.int32); b = np.zeros(N, np.int32)
        na = L.refm_search_by_projection(ncam, p(S['cam_off']), p(S['k']), p(S['desc']), p(S['bounds']), p(S['sf']), 8, p(S['occupied']), p(S['l2r']),
                                         p(S['r2l']), p(S['u_right']), S['nmp'], p(S['in_view']), p(S['px']), p(S['py']), p(S['lvl']), p(S['vc']),
                                         p(S['pxr']), p(S['depth']), p(S['bad']), p(S['has_obs']), p(S['md']), C.c_float(th), int(far),
                                         C.c_float(40.0), C.c_float(nnr), p(a))
        grids = (C.c_void_p * ncam)()
        for c in range(ncam):
            kc = S['k'][S['cam_off'][c]:S['cam_off'][c + 1]]
            grids[c] = Lo.orc_grid_create(p(np.ascontiguousarray(kc)), len(kc), C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
        nb = Lo.orc_search_by_projection(ncam, p(S['cam_off']), p(S['k']), p(S['desc']), grids, p(S['sf']), p(S['occupied']), p(S['l2r']), p(S['r2l']),
                                         p(S['u_right']), S['nmp'], p(S['in_view']), p(S['px']), p(S['py']), p(S['lvl']), p(S['vc']), p(S['pxr']),
                                         p(S['depth']), p(S['bad']), p(S['has_obs']), p(S['md']), C.c_float(th), int(far), C.c_float(40.0),
                                         C.c_float(nnr), p(b))
        for c in range(ncam):
            Lo.orc_grid_destroy(C.c_void_p(grids[c]))
        assert na == nb and na > 20, (trial, na, nb)
        assert np.array_equal(a, b), (trial, int((a != b).sum()))


def test_search_for_initialization():
    rng = np.random.default_rng(8)
    L = RM(); Lo = O.lib()
    for trial in range(20 * TRIALS_X):
        n1, n2 = int(rng.integers(100, 1500)), int(rng.integers(100, 1500))
        k2 = make_kps(n2, rng, lattice=6 if trial % 2 else None); k2['octave'][rng.random(n2) < 0.6] = 0
        d2 = synth_descriptors(n2, 900 + trial)
        src = rng.integers(0, n2, n1)
        k1 = k2[src].copy(); k1['x'] += rng.uniform(-6, 6, n1).astype(np.float32); k1['y'] += rng.uniform(-6, 6, n1).astype(np.float32)
        k1['angle'] = (k1['angle'] + rng.choice([0, 0, 0, 45, 200], n1)) % 360
        d1 = d2[src] ^ np.packbits(rng.random((n1, 256)) < 0.06, axis=1, bitorder='little')
        if trial % 3 == 0:
            d1[: n1 // 3] = d1[0]; d2[: n2 // 3] = d1[0]     # many equal distances: eviction / ownership bookkeeping
        bounds = np.array([0, 640, 0, 480], np.float32)
        prev = np.stack([k1['x'], k1['y']], 1).astype(np.float32)
        for chk in (0, 1):
            pa, pb = prev.copy(), prev.copy()
            a = np.zeros(n1, np.int32); b = np.zeros(n1, np.int32)
            na = L.refm_search_for_initialization(p(k1), p(d1), n1, p(k2), p(d2), n2, p(bounds), p(pa), 10 + 30 * (trial % 2), C.c_float(0.9), chk, p(a))
            g = Lo.orc_grid_create(p(k2), n2, C.c_float(0), C.c_float(640), C.c_float(0), C.c_float(480))
            nb = Lo.orc_search_for_initialization(p(k1), p(d1), n1, C.c_void_p(g), p(k2), p(d2), n2, p(pb), 10 + 30 * (trial % 2), C.c_float(0.9), chk, p(b))
            Lo.orc_grid_destroy(C.c_void_p(g))
            assert na == nb, (trial, chk, na, nb)
            assert np.array_equal(a, b) and np.array_equal(pa, pb), (trial, chk)


def _feature_vector(desc, nbits=5):
    node = desc[:, 0].astype(np.int32) & ((1 << nbits) - 1)
    order = np.argsort(node, kind='stable')
    nodes, counts = np.unique(node, return_counts=True)
    return nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)


PRED = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p)
# camera ids (KF1, KF2) of the carried pair states of orc_search_for_triangulation_full
STATE_CAMS = {0: (0, 0), 1: (0, 0), 2: (0, 1), 3: (1, 0), 4: (1, 1), 5: (0, 2), 6: (2, 0), 7: (2, 2), 8: (1, 3), 9: (3, 1), 10: (3, 3)}


@pytest.mark.parametrize('ncam', [1, 2, 4])
@pytest.mark.parametrize('coarse', [0, 1])
def test_search_for_triangulation(ncam, coarse):
    """SearchForTriangulation (src/ORBmatcher.cc:1131-1456) with every gate: bOnlyStereo, the epipole-distance reject of the monocular
    case, the camera-pair choice whose (pCamera1, pCamera2, R12, t12) is CARRIED across candidates when a pair has no branch, the
    last-minimum tie rule, the rotation histogram.  The reference's camera is the pixel camera of oracle/ref_matcher_driver.cpp whose
    epipolarConstrain is a hash of the keypoints and of the two camera ids, i.e. of the carried state."""
    rng = np.random.default_rng(40 + 10 * ncam + coarse)
    L = RM(); Lo = O.lib()
    L.refm_pair_hash.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    for trial in range(10 * TRIALS_X):
        per1 = [int(rng.integers(150, 500)) for _ in range(ncam)]; per2 = [int(rng.integers(150, 500)) for _ in range(ncam)]
        off1 = np.concatenate([[0], np.cumsum(per1)]).astype(np.int32); off2 = np.concatenate([[0], np.cumsum(per2)]).astype(np.int32)
        n1, n2 = int(off1[-1]), int(off2[-1])
        k2 = make_kps(n2, rng); d2 = synth_descriptors(n2, 4000 + trial)
        d2[10:60] = d2[10]                                   # exact duplicates -> last-minimum tie rule
        src = rng.integers(0, n2, n1)
        flips = np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        d1 = d2[src] ^ flips
        d1[::6] = synth_descriptors(len(d1[::6]), 43 + trial)
        k1 = make_kps(n1, rng)
        fv1, fv2 = _feature_vector(d1), _feature_vector(d2)
        fv1 = (fv1[0][::2].copy(), np.concatenate([[0], np.cumsum(np.diff(fv1[1])[::2])]).astype(np.int32),
               np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in range(0, len(fv1[0]), 2)]).astype(np.int32))   # some nodes only in KF2
        mp1 = (rng.random(n1) < 0.3).astype(np.uint8); mp2 = (rng.random(n2) < 0.3).astype(np.uint8)
        ur1 = ur2 = None; only_stereo = 0
        if ncam == 1 and trial % 2 == 1:
            ur1 = np.where(rng.random(n1) < 0.5, 10.0, -1.0).astype(np.float32); ur2 = np.where(rng.random(n2) < 0.5, 10.0, -1.0).astype(np.float32)
            only_stereo = int(trial % 4 == 3)
        ep = k2[['x', 'y']][int(rng.integers(0, n2))]; ep = np.array([ep['x'] + 2.0, ep['y'] - 1.0], np.float32)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        chk = trial % 2
        a = np.zeros(n1, np.int32); b = np.zeros(n1, np.int32)
        na = L.refm_search_for_triangulation(ncam, p(off1), p(k1), p(d1), p(mp1), p(ur1), len(fv1[0]), p(fv1[0]), p(fv1[1]), p(fv1[2]), p(off2), p(k2),
                                             p(d2), p(mp2), p(ur2), len(fv2[0]), p(fv2[0]), p(fv2[1]), p(fv2[2]), p(ep), p(sf), 8, only_stereo, coarse,
                                             chk, p(a))
        states = []

        def pred(i1, i2, state, user):
            states.append(state)
            c1, c2 = STATE_CAMS[state]
            return L.refm_pair_hash(k1['x'][i1], k1['y'][i1], k2['x'][i2], k2['y'][i2], c1, c2)
        cb = PRED(pred)
        nb = Lo.orc_search_for_triangulation_full(ncam, p(off1), p(k1), p(d1), p(mp1), p(ur1), len(fv1[0]), p(fv1[0]), p(fv1[1]), p(fv1[2]), p(off2),
                                                  p(k2), p(d2), p(mp2), p(ur2), len(fv2[0]), p(fv2[0]), p(fv2[1]), p(fv2[2]), p(ep), p(sf),
                                                  only_stereo, coarse, chk, cb, None, p(b))
        assert na == nb, (trial, na, nb)
        assert np.array_equal(a, b), (trial, int((a != b).sum()))
        if not only_stereo:
            assert na > 5
        if ncam == 4 and not coarse:
            assert len(set(states)) >= 8                      # the scenario exercises the carried camera-pair states


def sft_pinhole_scenarios(ncam, coarse):
    """SearchForTriangulation with real pinhole geometry: every KF2 keypoint is the image of a 3-D point in its camera; a KF1 keypoint
    planted on it is the image of the same point in ITS camera (+ up to 1.5 px of noise, 12 px for 40 % of them), so that the reference's own
    Pinhole::epipolarConstrain accepts most planted pairs under the right camera pair and few under a wrong (carried) one."""
    rng = np.random.default_rng(140 + 10 * ncam + coarse)
    cam = np.array([[400, 400, 320, 240], [410, 405, 318, 243], [395, 398, 322, 238], [402, 401, 321, 241]], np.float32)
    rel = np.array([[-0.12, 0.0, 0.0], [0.05, -0.03, 0.0], [-0.2, 0.03, 0.01]], np.float32)        # right, side-left, side-right from the left camera
    relc = np.concatenate([np.zeros((1, 3), np.float32), rel])
    for trial in range(8 * TRIALS_X):
        per1 = [int(rng.integers(150, 500)) for _ in range(ncam)]; per2 = [int(rng.integers(150, 500)) for _ in range(ncam)]
        off1 = np.concatenate([[0], np.cumsum(per1)]).astype(np.int32); off2 = np.concatenate([[0], np.cumsum(per2)]).astype(np.int32)
        n1, n2 = int(off1[-1]), int(off2[-1])
        k2 = make_kps(n2, rng); d2 = synth_descriptors(n2, 14000 + trial)
        d2[10:40] = d2[10]
        src = rng.integers(0, n2, n1)
        flips = np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        d1 = d2[src] ^ flips
        d1[::6] = synth_descriptors(len(d1[::6]), 143 + trial)
        k1 = make_kps(n1, rng)
        t1w = np.array([[-0.05, 0.02, -0.5], [-0.3, 0.02, 0.05]][trial % 2], np.float32)           # forward motion (epipole inside the image) / sideways
        c2 = np.searchsorted(off2, np.arange(n2), side='right') - 1; c1 = np.searchsorted(off1, np.arange(n1), side='right') - 1
        z = rng.uniform(2.0, 8.0, n2)
        X = np.stack([(k2['x'] - cam[c2, 2]) / cam[c2, 0] * z, (k2['y'] - cam[c2, 3]) / cam[c2, 1] * z, z], 1) - relc[c2]   # world = KF2's left camera
        Xc1 = X[src] + t1w + relc[c1]
        planted = (np.arange(n1) % 6 != 0) & (Xc1[:, 2] > 0.2)
        noise = np.where(rng.random(n1) < 0.6, 1.5, 12.0)                                          # 40 % are off the epipolar line by up to 12 px
        u1 = cam[c1, 0] * Xc1[:, 0] / Xc1[:, 2] + cam[c1, 2] + rng.uniform(-1, 1, n1) * noise
        v1 = cam[c1, 1] * Xc1[:, 1] / Xc1[:, 2] + cam[c1, 3] + rng.uniform(-1, 1, n1) * noise
        k1['x'][planted] = u1[planted].astype(np.float32); k1['y'][planted] = v1[planted].astype(np.float32)
        fv1, fv2 = _feature_vector(d1), _feature_vector(d2)
        mp1 = (rng.random(n1) < 0.2).astype(np.uint8); mp2 = (rng.random(n2) < 0.2).astype(np.uint8)
        ur1 = ur2 = None; only_stereo = 0
        if ncam == 1 and trial % 2 == 1:
            ur1 = np.where(rng.random(n1) < 0.5, 10.0, -1.0).astype(np.float32); ur2 = np.where(rng.random(n2) < 0.5, 10.0, -1.0).astype(np.float32)
            only_stereo = int(trial % 4 == 3)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        yield dict(trial=trial, ncam=ncam, coarse=coarse, off1=off1, off2=off2, n1=n1, n2=n2, k1=k1, k2=k2, d1=d1, d2=d2, fv1=fv1, fv2=fv2, mp1=mp1, mp2=mp2,
                   ur1=ur1, ur2=ur2, only_stereo=only_stereo, cam=np.ascontiguousarray(cam), t1w=t1w, rel=np.ascontiguousarray(rel), sf=sf,
                   sigma2=(sf * sf).astype(np.float32), chk=trial % 2)


def sft_row_on_uninitialised_pair(S):
    """KF1 index of the row that is tested on uninitialised geometry, or -1: if the FIRST candidate of the call that passes the descriptor
    gate is a camera combination without a branch in the pair selection (left / side-right, right / side-left, ...), the reference
    evaluates epipolarConstrain on R12 / t12 it never initialised (src/ORBmatcher.cc:1152-1154, 1301-1387) — undefined behaviour; that one
    row has nothing to be compared against (the product tests the candidate with the left-left pair)."""
    if S['ncam'] < 4:
        return -1
    branch = {(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0), (2, 2), (1, 3), (3, 1), (3, 3)}
    (nodes1, ptr1, ord1), (nodes2, ptr2, ord2) = S['fv1'], S['fv2']
    pos2 = {int(n): k for k, n in enumerate(nodes2)}
    for a, node in enumerate(nodes1):
        b = pos2.get(int(node))
        if b is None:
            continue
        for i1 in ord1[ptr1[a]:ptr1[a + 1]]:
            if S['mp1'][i1]:
                continue
            for i2 in ord2[ptr2[b]:ptr2[b + 1]]:
                if S['mp2'][i2] or int(np.unpackbits(S['d1'][i1] ^ S['d2'][i2]).sum()) > 50:
                    continue
                c1 = int(np.searchsorted(S['off1'], i1, side='right') - 1); c2 = int(np.searchsorted(S['off2'], i2, side='right') - 1)
                return -1 if (c1, c2) in branch else int(i1)
    return -1


def call_sft_pinhole(L, S):
    a = np.zeros(S['n1'], np.int32)
    fv1, fv2 = S['fv1'], S['fv2']
    n = L.refm_search_for_triangulation_pinhole(S['ncam'], p(S['off1']), p(S['k1']), p(S['d1']), p(S['mp1']), p(S['ur1']), len(fv1[0]), p(fv1[0]), p(fv1[1]),
                                                p(fv1[2]), p(S['off2']), p(S['k2']), p(S['d2']), p(S['mp2']), p(S['ur2']), len(fv2[0]), p(fv2[0]), p(fv2[1]),
                                                p(fv2[2]), p(S['cam']), p(S['t1w']), p(S['rel']), p(S['sf']), p(S['sigma2']), 8, S['only_stereo'], S['coarse'],
                                                S['chk'], p(a))
    return n, a


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_for_triangulation_pinhole_scenario(ncam):
    """(sanity of the pinhole scenario on the reference itself: the epipolar predicate — the reference's own Pinhole::epipolarConstrain —
    rejects a good part of what the coarse mode accepts, and still leaves matches)"""
    L = RM()
    tot = [0, 0]
    for Sf, Sc in zip(sft_pinhole_scenarios(ncam, 0), sft_pinhole_scenarios(ncam, 1)):
        nf, af = call_sft_pinhole(L, Sf)
        Sc2 = dict(Sf); Sc2['coarse'] = 1
        nc, ac = call_sft_pinhole(L, Sc2)
        assert nf == int((af >= 0).sum()) and nc == int((ac >= 0).sum())
        tot[0] += nf; tot[1] += nc
    assert tot[0] > 300 and tot[0] < 0.9 * tot[1], tot


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_bow_keyframe_frame(ncam):
    """SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), src/ORBmatcher.cc:349-666: per-camera best / second, the `|| true` ratio tests of
    cameras 1-3 nested in camera 0's TH_LOW test, already-matched rule, rotation histogram."""
    rng = np.random.default_rng(300 + ncam)
    L = RM(); Lo = O.lib()
    for trial in range(10 * TRIALS_X):
        nF = int(rng.integers(300, 1500)); nKF = int(rng.integers(300, 1500))
        dF = synth_descriptors(nF, 2900 + trial)
        dF[5:40] = dF[5]
        src = rng.integers(0, nF, nKF)
        flips = np.packbits(rng.random((nKF, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        dKF = dF[src] ^ flips
        dKF[::7] = synth_descriptors(len(dKF[::7]), 77 + trial)
        aF = rng.uniform(0, 360, nF).astype(np.float32)
        aKF = ((aF[src] + rng.choice([0, 0, 0, 30, 95, 190, 250], nKF)) % 360).astype(np.float32)
        fvK, fvF = _feature_vector(dKF), _feature_vector(dF)
        skip = (rng.random(nKF) < 0.25).astype(np.uint8)

        def split(n):
            cuts = np.sort(rng.integers(0, n + 1, ncam - 1)) if ncam > 1 else np.zeros(0, np.int64)
            return np.concatenate([[0], cuts, [n]]).astype(np.int32)
        offK, offF = split(nKF), split(nF)
        for chk in (0, 1):
            a = np.zeros(nF, np.int32); b = np.zeros(nF, np.int32)
            na = L.refm_search_by_bow_frame(len(fvK[0]), p(fvK[0]), p(fvK[1]), p(fvK[2]), len(fvF[0]), p(fvF[0]), p(fvF[1]), p(fvF[2]), p(dKF), nKF,
                                            p(dF), nF, p(skip), p(aKF), p(aF), ncam, p(offK), p(offF), C.c_float(0.7), chk, p(a))
            nb = Lo.orc_search_by_bow_frame(len(fvK[0]), p(fvK[0]), p(fvK[1]), p(fvK[2]), len(fvF[0]), p(fvF[0]), p(fvF[1]), p(fvF[2]), p(dKF), p(dF), nF,
                                            p(skip), p(aKF), p(aF), ncam, p(offF), C.c_float(0.7), chk, p(b))
            assert na == nb, (trial, chk, na, nb)
            assert np.array_equal(a, b), (trial, chk, int((a != b).sum()))


def bow_keyframes_scenarios():
    """(n1, n2, d1, d2, a1, a2, fv1, fv2, s1, s2) per trial: planted neighbours, duplicate rows, skipped features, rotated angles."""
    rng = np.random.default_rng(55)
    for trial in range(12 * TRIALS_X):
        n1, n2 = int(rng.integers(300, 1500)), int(rng.integers(300, 1500))
        d2 = synth_descriptors(n2, 5900 + trial)
        d2[5:40] = d2[5]
        src = rng.integers(0, n2, n1)
        flips = np.packbits(rng.random((n1, 256)) < 0.04, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        d1 = d2[src] ^ flips
        d1[::5] = synth_descriptors(len(d1[::5]), 99 + trial)
        a2 = rng.uniform(0, 360, n2).astype(np.float32)
        a1 = ((a2[src] + rng.choice([0, 0, 0, 30, 95, 190, 250], n1)) % 360).astype(np.float32)
        fv1, fv2 = _feature_vector(d1), _feature_vector(d2)
        s1 = (rng.random(n1) < 0.2).astype(np.uint8); s2 = (rng.random(n2) < 0.2).astype(np.uint8)
        yield trial, n1, n2, d1, d2, a1, a2, fv1, fv2, s1, s2


def call_bow_keyframes(L, S, chk):
    trial, n1, n2, d1, d2, a1, a2, fv1, fv2, s1, s2 = S
    a = np.zeros(n1, np.int32)
    na = L.refm_search_by_bow_keyframes(len(fv1[0]), p(fv1[0]), p(fv1[1]), p(fv1[2]), len(fv2[0]), p(fv2[0]), p(fv2[1]), p(fv2[2]), p(d1), n1,
                                        p(d2), n2, p(s1), p(s2), p(a1), p(a2), C.c_float(0.75), chk, p(a))
    return na, a


def test_search_by_bow_keyframe_keyframe():
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), src/ORBmatcher.cc:1006-1129: strict < TH_LOW, ratio test, vbMatched2, rotation filter."""
    L = RM(); Lo = O.lib()
    for S in bow_keyframes_scenarios():
        trial, n1, n2, d1, d2, a1, a2, fv1, fv2, s1, s2 = S
        for chk in (0, 1):
            b = np.zeros(n1, np.int32)
            na, a = call_bow_keyframes(L, S, chk)
            nb = Lo.orc_search_by_bow_keyframes(len(fv1[0]), p(fv1[0]), p(fv1[1]), p(fv1[2]), len(fv2[0]), p(fv2[0]), p(fv2[1]), p(fv2[2]), p(d1), n1,
                                                p(d2), n2, p(s1), p(s2), p(a1), p(a2), C.c_float(0.75), chk, p(b))
            assert na == nb and na > 10, (trial, chk, na, nb)
            assert np.array_equal(a, b), (trial, chk, int((a != b).sum()))


def sbp_last_scenarios(ncam):
    """One dict per trial for SearchByProjection(CurrentFrame, LastFrame): planted projections, some outside the image, competing points,
    occupied keypoints, rectified-stereo coordinates (one camera), the three direction gates, rotation filter on / off."""
    rng = np.random.default_rng(400 + ncam)
    w = h = 512
    for trial in range(12 * TRIALS_X):
        per = [int(rng.integers(200, 700)) for _ in range(ncam)]
        cam_off = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
        N = int(cam_off[-1])
        k = make_kps(N, rng, w, h, lattice=8 if trial % 2 else None)
        desc = synth_descriptors(N, int(rng.integers(1 << 30)))
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        npts = int(rng.integers(200, 1500))
        src = rng.integers(0, cam_off[1], npts)                          # planted on camera-0 keypoints
        px = (k['x'][src] + rng.uniform(-4, 4, npts)).astype(np.float32); py = (k['y'][src] + rng.uniform(-4, 4, npts)).astype(np.float32)
        out = rng.random(npts) < 0.08                                   # some projections fall outside the image
        px[out] = np.float32(-50.0)
        trl = np.array([rng.uniform(-30, 30), rng.uniform(-5, 5)], np.float32)
        has_mp = (rng.random(npts) < 0.9).astype(np.uint8)
        octv = np.clip(k['octave'][src] + rng.integers(-1, 2, npts), 0, 7).astype(np.int32)
        ang = ((k['angle'][src] + np.where(rng.random(npts) < 0.8, 20 + rng.uniform(-3, 3, npts), rng.uniform(0, 360, npts))) % 360).astype(np.float32)
        pd = synth_descriptors(npts, int(rng.integers(1 << 30)))
        own = rng.random(npts) < 0.75
        pd[own] = desc[src[own]] ^ np.packbits(rng.random((int(own.sum()), 256)) < 0.06, axis=1, bitorder='little')
        if trial % 3 == 0:
            pd[: npts // 4] = pd[0]
        has_obs = (rng.random(npts) < 0.85).astype(np.uint8)
        occupied = (rng.random(N) < 0.05).astype(np.uint8)
        occupied[rng.random(N) < 0.04] = 2                              # holds a point WITHOUT observations: claimable; NULL only if a claim is cleared
        mbf = np.float32(17.25)
        u_right = np.where(rng.random(N) < 0.4, k['x'] - rng.uniform(2, 40, N), -1).astype(np.float32) if ncam == 1 else None
        th = float([7.0, 15.0][trial % 2]); gate = trial % 3; chk = (trial // 3) % 2
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, ncam=ncam, cam_off=cam_off, N=N, k=k, desc=desc, sf=sf, npts=npts, px=px, py=py, trl=trl, has_mp=has_mp, octv=octv, ang=ang,
                   pd=pd, has_obs=has_obs, occupied=occupied, mbf=mbf, u_right=u_right, th=th, gate=gate, chk=chk, bounds=bounds, w=w, h=h)


def call_sbp_last(L, S):
    a = np.zeros(S['N'], np.int32)
    na = L.refm_search_by_projection_last(S['ncam'], p(S['cam_off']), p(S['k']), p(S['desc']), p(S['bounds']), p(S['sf']), 8, p(S['occupied']), p(S['u_right']),
                                          S['npts'], p(S['has_mp']), p(S['px']), p(S['py']), p(S['trl']), C.c_float(S['mbf']), p(S['octv']), p(S['ang']),
                                          p(S['has_obs']), p(S['pd']), C.c_float(S['th']), S['gate'], S['chk'], p(a))
    return na, a


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_projection_current_last(ncam):
    """SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), src/ORBmatcher.cc:1985-2413 — the tracking matcher:
    per-camera windows (the side cameras around the LEFT projection, D6), bForward / bBackward level gates, claims in point order, the
    empty-camera-0 `continue`, the rectified-stereo test, TH_HIGH, rotation histogram.  The reference computes the projections itself
    from poses and a camera: identity pose, pixel camera, right camera = left + translation."""
    L = RM(); Lo = O.lib()
    for S in sbp_last_scenarios(ncam):
        trial, cam_off, N, k, desc, sf, npts, px, py, trl = (S[x] for x in ('trial', 'cam_off', 'N', 'k', 'desc', 'sf', 'npts', 'px', 'py', 'trl'))
        has_mp, octv, ang, pd, has_obs, occupied, mbf, u_right, th, gate, chk, w, h = (S[x] for x in ('has_mp', 'octv', 'ang', 'pd', 'has_obs', 'occupied', 'mbf',
                                                                                                  'u_right', 'th', 'gate', 'chk', 'w', 'h'))
        na, a = call_sbp_last(L, S)
        # the same inputs as the array-level interface sees them
        valid = (has_mp.astype(bool) & (px >= 0) & (px <= w) & (py >= 0) & (py <= h)).astype(np.uint8)
        PX = np.stack([px, (px + trl[0]).astype(np.float32), px, px], 1).astype(np.float32)
        PY = np.stack([py, (py + trl[1]).astype(np.float32), py, py], 1).astype(np.float32)
        pur = (px - mbf).astype(np.float32)
        grids = (C.c_void_p * ncam)()
        for c in range(ncam):
            kc = np.ascontiguousarray(k[cam_off[c]:cam_off[c + 1]])
            grids[c] = Lo.orc_grid_create(p(kc), len(kc), C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
        b = np.zeros(N, np.int32)
        occ1 = (occupied == 1).astype(np.uint8)                          # what the array-level interface calls occupied
        nb = Lo.orc_search_by_projection_last(ncam, p(cam_off), p(k), p(desc), grids, p(sf), p(occ1), p(u_right), npts, p(valid), p(PX), p(PY), p(pur),
                                              p(octv), p(ang), p(has_obs), p(pd), C.c_float(th), gate, chk, 100, p(b))
        for c in range(ncam):
            Lo.orc_grid_destroy(C.c_void_p(grids[c]))
        assert na == nb and na > 20, (trial, na, nb)
        assert (occupied[a == -3] == 2).all()
        assert np.array_equal(np.where(a == -3, -1, a), b), (trial, int((a != b).sum()))


def test_fuse_keyframe_mappoints():
    """ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th, cameraID = 0), src/ORBmatcher.cc:1458-1640: the search loop — KeyFrame::GetFeaturesInArea
    window, level gate [l-1, l], mono (5.99) / stereo (7.8) reprojection gates, first minimum, TH_LOW — against the oracle's window search
    (orc_window_best), which is what the device entry point orbx_window_best is tested against."""
    rng = np.random.default_rng(66)
    L = RM(); Lo = O.lib()
    w, h = 640, 480
    for trial in range(12 * TRIALS_X):
        n = int(rng.integers(300, 2000)); nq = int(rng.integers(200, 1500))
        k = make_kps(n, rng, w, h, lattice=5 if trial % 2 else None)
        desc = synth_descriptors(n, 6900 + trial)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        inv_s2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
        src = rng.integers(0, n, nq)
        qx = (k['x'][src] + rng.uniform(-2.5, 2.5, nq)).astype(np.float32); qy = (k['y'][src] + rng.uniform(-2.5, 2.5, nq)).astype(np.float32)
        qx[rng.random(nq) < 0.05] = np.float32(-20.0)                        # outside the image: !IsInImage
        lvl = np.clip(k['octave'][src] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
        qd = desc[src] ^ np.packbits(rng.random((nq, 256)) < 0.07, axis=1, bitorder='little')
        qd[::6] = synth_descriptors(len(qd[::6]), 31 + trial)
        bf = np.float32(19.5)
        u_right = np.where(rng.random(n) < 0.5, k['x'] - bf + rng.uniform(-1.5, 1.5, n), -1).astype(np.float32) if trial % 2 == 0 else None
        valid = (rng.random(nq) < 0.95).astype(np.uint8)
        th = float([3.0, 5.0][trial % 2])
        bounds = np.array([0, w, 0, h], np.float32)
        a = np.zeros(nq, np.int32)
        nf = L.refm_fuse(p(k), p(desc), n, p(bounds), p(sf), p(inv_s2), 8, p(u_right), C.c_float(bf), nq, p(valid), p(qx), p(qy), p(lvl), p(qd),
                         C.c_float(th), p(a))
        # the same searches as the array-level interface sees them (the continues before the search are the caller's `valid`)
        v2 = (valid.astype(bool) & (qx >= 0) & (qx < w) & (qy >= 0) & (qy < h)).astype(np.uint8)
        radius = (np.float32(th) * sf[lvl]).astype(np.float32)
        q_ur = (qx - bf).astype(np.float32)
        g = Lo.orc_grid_create(p(k), n, C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
        bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32)
        Lo.orc_window_best(C.c_void_p(g), p(k), p(desc), None, p(u_right), p(inv_s2), nq, p(v2), p(qx), p(qy), p(q_ur), p(radius),
                           p((lvl - 1).astype(np.int32)), p(lvl), p(qd), p(bi), p(bd))
        Lo.orc_grid_destroy(C.c_void_p(g))
        exp = np.where(bd <= 50, bi, -1)
        assert nf == int((exp >= 0).sum()) and nf > 20, (trial, nf, int((exp >= 0).sum()))
        assert np.array_equal(a, exp), (trial, int((a != exp).sum()))


def fuse_graph_scenarios():
    """Fuse with every branch of the graph update: keyframe features holding a well-observed point / nothing / an unobserved point / a bad
    point, and map points that appear twice in the candidate list."""
    rng = np.random.default_rng(67)
    w, h = 640, 480
    for trial in range(10 * TRIALS_X):
        n = int(rng.integers(300, 2000)); nq = int(rng.integers(200, 1500))
        k = make_kps(n, rng, w, h, lattice=5 if trial % 2 else None)
        desc = synth_descriptors(n, 6700 + trial)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        inv_s2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
        src = rng.integers(0, n, nq)
        qx = (k['x'][src] + rng.uniform(-2.5, 2.5, nq)).astype(np.float32); qy = (k['y'][src] + rng.uniform(-2.5, 2.5, nq)).astype(np.float32)
        qx[rng.random(nq) < 0.05] = np.float32(-20.0)
        lvl = np.clip(k['octave'][src] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
        qd = desc[src] ^ np.packbits(rng.random((nq, 256)) < 0.07, axis=1, bitorder='little')
        qd[::6] = synth_descriptors(len(qd[::6]), 31 + trial)
        bf = np.float32(19.5)
        u_right = np.where(rng.random(n) < 0.5, k['x'] - bf + rng.uniform(-1.5, 1.5, n), -1).astype(np.float32) if trial % 2 == 0 else None
        valid = (rng.random(nq) < 0.95).astype(np.uint8)
        kf_state = rng.choice([0, 1, 2, 3], n, p=[0.4, 0.3, 0.2, 0.1]).astype(np.uint8)
        dup_of = np.full(nq, -1, np.int32)
        for q in rng.choice(np.arange(1, nq), nq // 10, replace=False):
            dup_of[q] = int(rng.integers(0, q))
        th = float([3.0, 5.0][trial % 2])
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, n=n, nq=nq, k=k, desc=desc, sf=sf, inv_s2=inv_s2, qx=qx, qy=qy, lvl=lvl, qd=qd, bf=bf, u_right=u_right, valid=valid,
                   kf_state=kf_state, dup_of=dup_of, th=th, bounds=bounds)


def call_fuse_graph(L, S):
    q_repl = np.zeros(S['nq'], np.int32); kf_mp = np.zeros(S['n'], np.int32); holder_repl = np.zeros(S['n'], np.int32)
    nf = L.refm_fuse_graph(p(S['k']), p(S['desc']), S['n'], p(S['bounds']), p(S['sf']), p(S['inv_s2']), 8, p(S['u_right']), C.c_float(S['bf']), p(S['kf_state']),
                           S['nq'], p(S['valid']), p(S['dup_of']), p(S['qx']), p(S['qy']), p(S['lvl']), p(S['qd']), C.c_float(S['th']), p(q_repl), p(kf_mp),
                           p(holder_repl))
    return nf, q_repl, kf_mp, holder_repl


def test_fuse_graph_updates_reference_runs_every_branch():
    """(sanity of the scenario generator on the reference itself: each of the four outcomes occurs)"""
    L = RM()
    seen = np.zeros(3, np.int64)
    for S in fuse_graph_scenarios():
        nf, q_repl, kf_mp, holder_repl = call_fuse_graph(L, S)
        assert nf > 20
        assert ((kf_mp >= 0) <= (S['kf_state'] == 1)).all() and ((holder_repl >= 0) <= (S['kf_state'] == 2)).all()
        assert (S['kf_state'][q_repl[q_repl >= 0]] == 0).all() and (q_repl <= -2).any()
        seen += [(q_repl >= 0).sum(), (kf_mp >= 0).sum(), (holder_repl >= 0).sum()]
    assert (seen > 50).all(), seen


def _oracle_last(Lo, k, desc, w, h, sf, occupied, npts, valid, px, py, octv, ang, pd, th, gate, chk, max_dist):
    """orc_search_by_projection_last for one camera block (what orbx_search_by_projection_last is tested against on the GPU)."""
    n = len(k)
    cam_off = np.array([0, n], np.int32)
    grids = (C.c_void_p * 1)()
    grids[0] = Lo.orc_grid_create(p(k), n, C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
    PX = np.repeat(px[:, None], 4, 1).astype(np.float32); PY = np.repeat(py[:, None], 4, 1).astype(np.float32)
    b = np.zeros(n, np.int32)
    nb = Lo.orc_search_by_projection_last(1, p(cam_off), p(k), p(desc), grids, p(sf), p(occupied), None, npts, p(valid), p(PX), p(PY), None, p(octv),
                                          p(ang), None, p(pd), C.c_float(th), gate, chk, max_dist, p(b))
    Lo.orc_grid_destroy(C.c_void_p(grids[0]))
    return nb, b


def _planted_points(rng, k, desc, npts, w, h, spread=4.0):
    src = rng.integers(0, len(k), npts)
    px = (k['x'][src] + rng.uniform(-spread, spread, npts)).astype(np.float32); py = (k['y'][src] + rng.uniform(-spread, spread, npts)).astype(np.float32)
    px[rng.random(npts) < 0.06] = np.float32(-30.0)                           # outside the image
    lvl = np.clip(k['octave'][src] + rng.integers(-1, 2, npts), 0, 7).astype(np.int32)
    pd = synth_descriptors(npts, int(rng.integers(1 << 30)))
    own = rng.random(npts) < 0.75
    pd[own] = desc[src[own]] ^ np.packbits(rng.random((int(own.sum()), 256)) < 0.06, axis=1, bitorder='little')
    return src, px, py, lvl, pd


def reloc_scenarios():
    rng = np.random.default_rng(77)
    w = h = 512
    for trial in range(12 * TRIALS_X):
        n = int(rng.integers(300, 1500)); npts = int(rng.integers(200, 1500))
        k = make_kps(n, rng, w, h, lattice=8 if trial % 2 else None)
        desc = synth_descriptors(n, 7700 + trial)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        src, px, py, lvl, pd = _planted_points(rng, k, desc, npts, w, h)
        if trial % 3 == 0:
            pd[: npts // 4] = pd[0]                                           # many points compete for the same keypoints
        ang = ((k['angle'][src] + np.where(rng.random(npts) < 0.8, 20 + rng.uniform(-3, 3, npts), rng.uniform(0, 360, npts))) % 360).astype(np.float32)
        has_mp = (rng.random(npts) < 0.9).astype(np.uint8); found = (rng.random(npts) < 0.1).astype(np.uint8)
        occupied = (rng.random(n) < 0.05).astype(np.uint8)
        th = float([10.0, 15.0][trial % 2]); orb_dist = [100, 64][trial % 2]; chk = (trial // 2) % 2
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, n=n, npts=npts, k=k, desc=desc, sf=sf, px=px, py=py, lvl=lvl, pd=pd, ang=ang, has_mp=has_mp, found=found, occupied=occupied,
                   th=th, orb_dist=orb_dist, chk=chk, bounds=bounds, w=w, h=h)


def call_reloc(L, S):
    a = np.zeros(S['n'], np.int32)
    na = L.refm_search_by_projection_reloc(p(S['k']), p(S['desc']), S['n'], p(S['bounds']), p(S['sf']), 8, p(S['occupied']), S['npts'], p(S['has_mp']),
                                           p(S['found']), p(S['px']), p(S['py']), p(S['lvl']), p(S['ang']), p(S['pd']), C.c_float(S['th']), S['orb_dist'],
                                           S['chk'], p(a))
    return na, a


def test_search_by_projection_relocalisation():
    """SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2415-2535: inclusive image
    bounds, level gate [l-1, l+1] inside GetFeaturesInArea, claims in point order, ORBdist, rotation histogram — against
    orc_search_by_projection_last (gate 0), the function the device entry point is tested against."""
    L = RM(); Lo = O.lib()
    for S in reloc_scenarios():
        na, a = call_reloc(L, S)
        px, py, w, h = S['px'], S['py'], S['w'], S['h']
        valid = (S['has_mp'].astype(bool) & ~S['found'].astype(bool) & (px >= 0) & (px <= w) & (py >= 0) & (py <= h)).astype(np.uint8)
        nb, b = _oracle_last(Lo, S['k'], S['desc'], w, h, S['sf'], S['occupied'], S['npts'], valid, px, py, S['lvl'], S['ang'], S['pd'], S['th'], 0, S['chk'],
                             S['orb_dist'])
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))


def kf_sim3_scenarios(with_kfs):
    rng = np.random.default_rng(88 + with_kfs)
    w, h = 640, 480
    for trial in range(10 * TRIALS_X):
        n = int(rng.integers(300, 1500)); npts = int(rng.integers(200, 1500))
        k = make_kps(n, rng, w, h, lattice=8 if trial % 2 else None)
        desc = synth_descriptors(n, 8800 + trial)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        src, px, py, lvl, pd = _planted_points(rng, k, desc, npts, w, h)
        if trial % 3 == 0:
            pd[: npts // 4] = pd[0]
        matched_in = (rng.random(n) < 0.08).astype(np.uint8)
        skip = rng.choice([0, 1, 2], npts, p=[0.9, 0.06, 0.04]).astype(np.uint8)
        skip[np.cumsum(skip == 2) > int(matched_in.sum())] = 0                # every 'already found' point needs an occupied keypoint to sit in
        th = [4, 8][trial % 2]; ratio = [1.0, 0.8, 0.9][trial % 3]
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, n=n, npts=npts, k=k, desc=desc, sf=sf, px=px, py=py, lvl=lvl, pd=pd, matched_in=matched_in, skip=skip, th=th, ratio=ratio,
                   bounds=bounds, w=w, h=h, with_kfs=with_kfs)


def call_kf_sim3(L, S):
    a = np.zeros(S['n'], np.int32); ok = C.c_int(0)
    na = L.refm_search_by_projection_sim3(p(S['k']), p(S['desc']), S['n'], p(S['bounds']), p(S['sf']), 8, p(S['matched_in']), S['npts'], p(S['skip']), p(S['px']),
                                          p(S['py']), p(S['lvl']), p(S['pd']), S['th'], C.c_float(S['ratio']), S['with_kfs'], p(a), C.byref(ok))
    return na, a, ok.value


@pytest.mark.parametrize('with_kfs', [0, 1])
def test_search_by_projection_keyframe_sim3(with_kfs):
    """Both SearchByProjection(KeyFrame*, Sim3f&, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming, cameraID) overloads,
    src/ORBmatcher.cc:668-893: IsInImage (half-open), KeyFrame::GetFeaturesInArea without level gate, the [l-1, l] gate in the loop,
    vpMatched as occupancy, TH_LOW * ratioHamming — against orc_search_by_projection_last (gate 3)."""
    L = RM(); Lo = O.lib()
    for S in kf_sim3_scenarios(with_kfs):
        na, a, ok = call_kf_sim3(L, S)
        px, py, w, h = S['px'], S['py'], S['w'], S['h']
        valid = ((S['skip'] == 0) & (px >= 0) & (px < w) & (py >= 0) & (py < h)).astype(np.uint8)
        max_dist = int(np.floor(np.float32(50) * np.float32(S['ratio'])))
        ang = np.zeros(S['npts'], np.float32)
        nb, b = _oracle_last(Lo, S['k'], S['desc'], w, h, S['sf'], S['matched_in'], S['npts'], valid, px, py, S['lvl'], ang, S['pd'], float(S['th']), 3, 0, max_dist)
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))
        assert not with_kfs or ok == 1


def sim3_scenarios():
    rng = np.random.default_rng(99)
    w, h = 640, 480
    for trial in range(10 * TRIALS_X):
        n1 = int(rng.integers(300, 1200)); n2 = int(rng.integers(300, 1200))
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        k2 = make_kps(n2, rng, w, h, lattice=6 if trial % 2 else None); d2 = synth_descriptors(n2, 9900 + trial)
        # KF1's keypoints: most of them are KF2 keypoints seen again (so that the two searches can agree)
        src = rng.integers(0, n2, n1)
        k1 = make_kps(n1, rng, w, h)
        again = rng.random(n1) < 0.7
        k1['x'][again] = (k2['x'][src[again]] + rng.uniform(-2, 2, int(again.sum()))).astype(np.float32)
        k1['y'][again] = (k2['y'][src[again]] + rng.uniform(-2, 2, int(again.sum()))).astype(np.float32)
        k1['octave'][again] = k2['octave'][src[again]]
        d1 = synth_descriptors(n1, 9950 + trial)
        d1[again] = d2[src[again]] ^ np.packbits(rng.random((int(again.sum()), 256)) < 0.05, axis=1, bitorder='little')
        has1 = (rng.random(n1) < 0.85).astype(np.uint8); has2 = (rng.random(n2) < 0.85).astype(np.uint8)
        # a map point of KF1 projects near its own keypoint in KF2's image (identity Sim3), with its keypoint's descriptor
        u1 = (k1['x'] + rng.uniform(-2, 2, n1)).astype(np.float32); v1 = (k1['y'] + rng.uniform(-2, 2, n1)).astype(np.float32)
        u2 = (k2['x'] + rng.uniform(-2, 2, n2)).astype(np.float32); v2 = (k2['y'] + rng.uniform(-2, 2, n2)).astype(np.float32)
        u1[rng.random(n1) < 0.05] = np.float32(-20.0); v2[rng.random(n2) < 0.05] = np.float32(h + 5.0)
        l1 = np.clip(k1['octave'] + rng.integers(0, 2, n1), 0, 7).astype(np.int32); l2 = np.clip(k2['octave'] + rng.integers(0, 2, n2), 0, 7).astype(np.int32)
        m1 = d1.copy(); m2 = d2.copy()
        pre12 = np.where(rng.random(n1) < 0.06, rng.integers(-2, n2 + 3, n1), -1).astype(np.int32)   # -2: not in KF2; >= n2: ignored (:1802)
        pre12[pre12 == -1] = -1
        th = float([7.5, 10.0][trial % 2])
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, n1=n1, n2=n2, sf=sf, k1=k1, k2=k2, d1=d1, d2=d2, has1=has1, has2=has2, u1=u1, v1=v1, u2=u2, v2=v2, l1=l1, l2=l2, m1=m1, m2=m2,
                   pre12=pre12, th=th, bounds=bounds, w=w, h=h)


def call_sim3(L, S):
    a = np.zeros(S['n1'], np.int32)
    na = L.refm_search_by_sim3(p(S['k1']), p(S['d1']), S['n1'], p(S['k2']), p(S['d2']), S['n2'], p(S['bounds']), p(S['sf']), 8, p(S['has1']), p(S['u1']),
                               p(S['v1']), p(S['l1']), p(S['m1']), p(S['has2']), p(S['u2']), p(S['v2']), p(S['l2']), p(S['m2']), p(S['pre12']),
                               C.c_float(S['th']), p(a))
    return na, a


def test_search_by_sim3():
    """ORBmatcher::SearchBySim3, src/ORBmatcher.cc:1771-1983: the already-matched bookkeeping through GetIndexInKeyFrame, the two window
    searches (TH_HIGH, [l-1, l] gate, first minimum) and the agreement check — against orc_search_by_sim3."""
    L = RM(); Lo = O.lib()
    for S in sim3_scenarios():
        na, a = call_sim3(L, S)
        n1, n2, w, h, sf, th = S['n1'], S['n2'], S['w'], S['h'], S['sf'], S['th']
        k1, k2, d1, d2, u1, v1, u2, v2, l1, l2, m1, m2, pre12 = (S[x] for x in ('k1', 'k2', 'd1', 'd2', 'u1', 'v1', 'u2', 'v2', 'l1', 'l2', 'm1', 'm2', 'pre12'))
        already1 = pre12 != -1                                                 # vpMatches12[i] != NULL
        already2 = np.zeros(n2, bool)
        idx2 = pre12[(pre12 >= 0) & (pre12 < n2)]
        already2[idx2] = True
        val1 = (S['has1'].astype(bool) & ~already1 & (u1 >= 0) & (u1 < w) & (v1 >= 0) & (v1 < h)).astype(np.uint8)
        val2 = (S['has2'].astype(bool) & ~already2 & (u2 >= 0) & (u2 < w) & (v2 >= 0) & (v2 < h)).astype(np.uint8)
        r1 = (np.float32(th) * sf[l1]).astype(np.float32); r2 = (np.float32(th) * sf[l2]).astype(np.float32)
        g1 = Lo.orc_grid_create(p(k1), n1, C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
        g2 = Lo.orc_grid_create(p(k2), n2, C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
        b = np.zeros(n1, np.int32)
        nb = Lo.orc_search_by_sim3(C.c_void_p(g1), p(k1), p(d1), n1, C.c_void_p(g2), p(k2), p(d2), n2, p(val1), p(u1), p(v1), p(r1), p(l1), p(m1),
                                   p(val2), p(u2), p(v2), p(r2), p(l2), p(m2), p(b))
        Lo.orc_grid_destroy(C.c_void_p(g1)); Lo.orc_grid_destroy(C.c_void_p(g2))
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))


def fuse_sim3_scenarios():
    rng = np.random.default_rng(111)
    w, h = 640, 480
    for trial in range(10 * TRIALS_X):
        n = int(rng.integers(300, 2000)); nq = int(rng.integers(200, 1500))
        k = make_kps(n, rng, w, h, lattice=5 if trial % 2 else None)
        desc = synth_descriptors(n, 11100 + trial)
        sf = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
        src, qx, qy, lvl, qd = _planted_points(rng, k, desc, nq, w, h, spread=2.5)
        lvl = np.clip(k['octave'][src] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
        has_mp = (rng.random(n) < 0.7).astype(np.uint8)
        valid = (rng.random(nq) < 0.95).astype(np.uint8)
        th = float([3.0, 5.0][trial % 2])
        bounds = np.array([0, w, 0, h], np.float32)
        yield dict(trial=trial, n=n, nq=nq, k=k, desc=desc, sf=sf, qx=qx, qy=qy, lvl=lvl, qd=qd, has_mp=has_mp, valid=valid, th=th, bounds=bounds, w=w, h=h)


def call_fuse_sim3(L, S):
    a = np.zeros(S['nq'], np.int32)
    nf = L.refm_fuse_sim3(p(S['k']), p(S['desc']), S['n'], p(S['bounds']), p(S['sf']), 8, p(S['has_mp']), S['nq'], p(S['valid']), p(S['qx']), p(S['qy']),
                          p(S['lvl']), p(S['qd']), C.c_float(S['th']), p(a))
    return nf, a


def test_fuse_keyframe_sim3():
    """ORBmatcher::Fuse(KeyFrame*, Sim3f &Scw, vpPoints, th, vpReplacePoint), src/ORBmatcher.cc:1648-1769, the loop-closing fuse: the search
    loop without reprojection gates ([l-1, l] level gate, first minimum, TH_LOW) — against orc_window_best."""
    L = RM(); Lo = O.lib()
    for S in fuse_sim3_scenarios():
        nf, a = call_fuse_sim3(L, S)
        n, nq, k, desc, sf, qx, qy, lvl, qd, w, h = (S[x] for x in ('n', 'nq', 'k', 'desc', 'sf', 'qx', 'qy', 'lvl', 'qd', 'w', 'h'))
        v2 = (S['valid'].astype(bool) & (qx >= 0) & (qx < w) & (qy >= 0) & (qy < h)).astype(np.uint8)
        radius = (np.float32(S['th']) * sf[lvl]).astype(np.float32)
        g = Lo.orc_grid_create(p(k), n, C.c_float(0), C.c_float(w), C.c_float(0), C.c_float(h))
        bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32)
        Lo.orc_window_best(C.c_void_p(g), p(k), p(desc), None, None, None, nq, p(v2), p(qx), p(qy), None, p(radius), p((lvl - 1).astype(np.int32)),
                           p(lvl), p(qd), p(bi), p(bd))
        Lo.orc_grid_destroy(C.c_void_p(g))
        exp = np.where(bd <= 50, bi, -1)
        assert nf == int((exp >= 0).sum()) and nf > 20, (S['trial'], nf, int((exp >= 0).sum()))
        assert np.array_equal(a, exp), (S['trial'], int((a != exp).sum()))
