"""Pins the CPU oracle against OpenCV itself (cv2 wheel) — the third-party library the reference
calls for every arithmetic primitive on the path:
  cv::resize            ORBextractor.cc:1083      cv::copyMakeBorder  :1085-1101
  cv::FAST              :745-769                  cv::GaussianBlur    :1036
  cv::fastAtan2         :42                       cv::BFMatcher       Frame.cc:1427,1483
and, end to end, against an independent Python restatement of operator() built from cv2 primitives.
"""
import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_image, synth_descriptors

cv2 = pytest.importorskip('cv2')

GEOMS = [(752, 480), (720, 540), (512, 512), (1280, 1024), (131, 97)]


def pyramid_sizes(w, h, nlevels=8, sf=1.2):
    t = O.OracleExtractor(1000, sf, nlevels, 20, 7).tables()
    return [(int(np.rint(np.float32(w) * t['inv_sf'][l])), int(np.rint(np.float32(h) * t['inv_sf'][l])))
            for l in range(nlevels)]


def test_tables_match_survey():
    t = O.OracleExtractor(1000, 1.2, 8, 20, 7).tables()
    assert list(t['nfeat']) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(O.OracleExtractor(1500, 1.2, 8, 20, 7).tables()['nfeat']) == [326, 271, 226, 189, 157, 131, 109, 91]
    assert list(O.OracleExtractor(2000, 1.2, 8, 20, 7).tables()['nfeat']) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(O.OracleExtractor(500, 1.2, 8, 20, 7).tables()['nfeat']) == [109, 90, 75, 63, 52, 44, 36, 31]
    assert list(t['umax']) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert pyramid_sizes(752, 480) == [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193),
                                       (252, 161), (210, 134)]
    assert pyramid_sizes(1280, 1024)[-1] == (357, 286)


@pytest.mark.parametrize('w,h', GEOMS)
def test_resize_chain_bitexact(w, h):
    rng = np.random.default_rng(w * 7 + h)
    cur = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    for (dw, dh) in pyramid_sizes(w, h)[1:]:
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = O.resize_linear(cur, dw, dh)
        assert np.array_equal(ref, got)
        cur = ref


@pytest.mark.parametrize('w,h', [(64, 48), (210, 134), (20, 25)])
def test_border_bitexact(w, h):
    img = np.random.default_rng(3).integers(0, 256, size=(h, w), dtype=np.uint8)
    ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, O.border101(img, 19))


@pytest.mark.parametrize('seed', range(6))
@pytest.mark.parametrize('thr', [20, 15, 7])
def test_fast_bitexact_incl_order(seed, thr):
    rng = np.random.default_rng(seed)
    if seed % 3 == 0:
        img = rng.integers(0, 256, size=(47, 53), dtype=np.uint8)
    elif seed % 3 == 1:
        img = synth_image(96, 80, seed, flat_region=False)[10:51, 20:61]          # a 41x41 cell view
    else:
        img = (rng.integers(0, 4, size=(41, 44)) * 80).astype(np.uint8)           # blocky, many score ties
    det = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = det.detect(np.ascontiguousarray(img))
    xs, ys, sc = O.fast(img, thr)
    assert len(ref) == len(xs)
    assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in ref] == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))


def test_fast_tiny_and_flat():
    assert len(O.fast(np.zeros((6, 40), np.uint8), 7)[0]) == 0
    assert len(O.fast(np.full((41, 41), 255, np.uint8), 7)[0]) == 0
    det = cv2.FastFeatureDetector_create(7, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    img = np.random.default_rng(0).integers(0, 256, size=(7, 7), dtype=np.uint8)
    assert len(det.detect(img)) == len(O.fast(img, 7)[0])


@pytest.mark.parametrize('w,h', [(752, 480), (347, 260), (143, 143), (31, 9)])
@pytest.mark.parametrize('kind', ['random', 'smooth', 'const'])
def test_blur_bitexact(w, h, kind):
    if kind == 'random':
        img = np.random.default_rng(1).integers(0, 256, size=(h, w), dtype=np.uint8)
    elif kind == 'smooth':
        img = synth_image(w, h, 5)
    else:
        img = np.full((h, w), 255, np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, O.blur7(img))


def test_fast_atan2_bitexact():
    rng = np.random.default_rng(11)
    y = rng.integers(-2400000, 2400000, size=20000).astype(np.float32)
    x = rng.integers(-2400000, 2400000, size=20000).astype(np.float32)
    y[:4] = [0, 0, 5, -5]; x[:4] = [0, -5, 0, 0]
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    got = O.fast_atan2(y, x)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    assert got[0] == 0.0 and got[1] == 180.0


def test_knn2_matches_bfmatcher_order():
    rng = np.random.default_rng(2)
    t = synth_descriptors(300, 1)
    t[:100, 4:] = 0        # tie-heavy: many equal distances
    q = np.concatenate([t[rng.integers(0, 300, 150)], synth_descriptors(50, 3)])
    q[:, :2] ^= rng.integers(0, 4, size=(200, 2), dtype=np.uint8)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
    idx, dist = O.knn2(q, t)
    for i, m in enumerate(ref):
        assert [mm.trainIdx for mm in m] == idx[i].tolist()
        assert [int(mm.distance) for mm in m] == dist[i].tolist()
    # one-row train set: a single neighbour
    idx1, dist1 = O.knn2(q[:5], t[:1])
    assert (idx1[:, 1] == -1).all() and (idx1[:, 0] == 0).all()


def test_hamming_is_popcount():
    a = synth_descriptors(64, 7); b = synth_descriptors(64, 8)
    for i in range(64):
        assert O.hamming(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


# ---- independent Python/cv2 restatement of operator() (ORBextractor.cc:987-1071) -------------------------
def py_extract(img, nf, ini, mn, lap, sel_from_oracle):
    """Glue in Python, primitives from cv2.  DistributeOctTree's node order is taken from the oracle
    (sel_from_oracle: per level candidate -> selection), everything else is independent."""
    t = O.OracleExtractor(nf, 1.2, 8, ini, mn).tables()
    h, w = img.shape
    levels = []
    for l in range(8):
        if l == 0:
            cur = img
        else:
            dw = int(np.rint(np.float32(w) * t['inv_sf'][l])); dh = int(np.rint(np.float32(h) * t['inv_sf'][l]))
            cur = cv2.resize(levels[-1], (dw, dh), interpolation=cv2.INTER_LINEAR)
        levels.append(cur)
    cands = []
    for l, L in enumerate(levels):
        hh, ww = L.shape
        minB, maxBX, maxBY = 16, ww - 16, hh - 16
        width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
        nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
        wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
        out = []
        for i in range(nRows):
            iniY = minB + i * hCell; maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell; maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                cell = np.ascontiguousarray(L[iniY:maxY, iniX:maxX])
                ks = cv2.FastFeatureDetector_create(ini, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(cell)
                if not ks:
                    ks = cv2.FastFeatureDetector_create(mn, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(cell)
                out += [(int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)) for k in ks]
        cands.append(out)
    return levels, cands


@pytest.mark.parametrize('w,h,nf,ini', [(752, 480, 1000, 20), (512, 512, 1500, 20)])
def test_pipeline_vs_python_cv2_restatement(w, h, nf, ini):
    img = synth_image(w, h, 1000 + w)
    E = O.OracleExtractor(nf, 1.2, 8, ini, 7)
    mono, kps, desc = E(img, (0, 0))
    levels, cands = py_extract(img, nf, ini, 7, (0, 0), None)
    for l in range(8):
        assert np.array_equal(E.level(l)[19:-19, 19:-19], levels[l])
        assert np.array_equal(E.level(l), cv2.copyMakeBorder(levels[l], 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
        xs, ys, sc = E.level_candidates(l)
        assert list(zip(xs.tolist(), ys.tolist(), sc.tolist())) == cands[l]
        kl = E.level_keypoints(l)
        if len(kl):
            blur = cv2.GaussianBlur(levels[l].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
            assert np.array_equal(E.level_blur(l), blur)
    assert mono == len(kps) and len(kps) >= nf * 0.9
    # orientation vs a direct numpy moment + cv2.fastAtan2
    um = E.tables()['umax']
    for l in (0, 3, 7):
        kl = E.level_keypoints(l)
        L = levels[l].astype(np.int64)
        for k in kl[:40]:
            x, y = int(k['x']), int(k['y'])
            m10 = m01 = 0
            for v in range(-15, 16):
                d = um[abs(v)]
                row = L[y + v, x - d:x + d + 1]
                m10 += int((np.arange(-d, d + 1) * row).sum()); m01 += v * int(row.sum())
            assert np.float32(cv2.fastAtan2(float(m01), float(m10))) == k['angle']
